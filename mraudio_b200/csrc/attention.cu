// Fused multi-head attention core for the Q-Former (head_dim 64): O = softmax(Q K^T / 8 + mask) V per (row, head).
//
// The problems are tiny (self: 64x64 keys; cross: 32 queries x 257 keys) and there are thousands of them, so this
// is a flash-style register kernel: one CTA per (row, head); each warp owns 16 query rows; keys are streamed through
// shared memory in chunks of 64 with an online softmax, so the [rows, heads, Sq, Sk] score tensor the reference
// materialises (HF port modeling_instructblip.py:512-536) never exists in HBM.  Tensor-core work uses
// mma.sync.m16n8k16 bf16 (fp32 accumulate): at 1.6 % of the path's FLOPs these tiles are too small for tcgen05's
// 128-row atoms to pay off.
#include "common.h"
#include "ptx.cuh"

namespace mra {
namespace {

constexpr int HD = 64;        // head dim
constexpr int KC = 64;        // keys per chunk
constexpr int LDS_ROW = 72;   // padded smem row (bf16 elements): 144 B stride -> conflict-free fragment loads
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    const __nv_bfloat16* q; int64_t ldq;
    const __nv_bfloat16* k; int64_t ldk;
    const __nv_bfloat16* v; int64_t ldv;
    __nv_bfloat16* o; int64_t ldo;
    const float* add_mask;
    int rows, heads, Sq, Sk, nq_split, kv_dense;
};

__device__ __forceinline__ int64_t split_index(int r, int i, int rows, int nsplit, int S) {
    return i < nsplit ? static_cast<int64_t>(r) * nsplit + i
                      : static_cast<int64_t>(rows) * nsplit + static_cast<int64_t>(r) * (S - nsplit) + (i - nsplit);
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(ptx::smem_u32(smem_row_ptr)));
}

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attention_kernel(const Params p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int sq_pad = (p.Sq + 15) & ~15;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [sq_pad][72]
    __nv_bfloat16* sK = sQ + sq_pad * LDS_ROW;                        // [64][72]
    __nv_bfloat16* sV = sK + KC * LDS_ROW;                            // [64][72]  (row = key, col = dim)
    float* sM = reinterpret_cast<float*>(sV + KC * LDS_ROW);          // [64] additive mask * log2e (or -inf)

    const int head = blockIdx.x % p.heads;
    const int r = blockIdx.x / p.heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = NWARPS * 32;

    // ---- stage Q (all Sq rows of this (row, head)); rows >= Sq are zero
    for (int idx = tid; idx < sq_pad * 8; idx += NT) {
        const int i = idx >> 3, c = idx & 7;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (i < p.Sq) {
            const int64_t gi = split_index(r, i, p.rows, p.nq_split, p.Sq);
            val = *reinterpret_cast<const uint4*>(p.q + gi * p.ldq + head * HD + c * 8);
        }
        *reinterpret_cast<uint4*>(sQ + i * LDS_ROW + c * 8) = val;
    }

    const int nrb = sq_pad / 16;
    const int nchunks = (p.Sk + KC - 1) / KC;
    const float scale_log2 = 0.125f * LOG2E;

    for (int rb0 = 0; rb0 < nrb; rb0 += NWARPS) {
        const int rb = rb0 + warp;
        const bool active = rb < nrb;
        uint32_t qf[4][4];
        float o_acc[8][4];
        float m_run[2] = {-INFINITY, -INFINITY};
        float l_run[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }

        for (int kc = 0; kc < nchunks; ++kc) {
            __syncthreads();  // previous chunk fully consumed (and sQ staged, first time)
            if (kc == 0 && active) {
                const __nv_bfloat16* qb = sQ + (rb * 16) * LDS_ROW;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    qf[ks][0] = *reinterpret_cast<const uint32_t*>(qb + g * LDS_ROW + ks * 16 + 2 * t);
                    qf[ks][1] = *reinterpret_cast<const uint32_t*>(qb + (g + 8) * LDS_ROW + ks * 16 + 2 * t);
                    qf[ks][2] = *reinterpret_cast<const uint32_t*>(qb + g * LDS_ROW + ks * 16 + 2 * t + 8);
                    qf[ks][3] = *reinterpret_cast<const uint32_t*>(qb + (g + 8) * LDS_ROW + ks * 16 + 2 * t + 8);
                }
            }
            // ---- load K / V chunk (keys kc*64 ..), zero-fill past Sk; build the mask row
            for (int idx = tid; idx < KC * 8; idx += NT) {
                const int j = idx >> 3, c = idx & 7;
                const int key = kc * KC + j;
                uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
                if (key < p.Sk) {
                    const int64_t gi = p.kv_dense ? static_cast<int64_t>(r) * p.Sk + key
                                                  : split_index(r, key, p.rows, p.nq_split, p.Sk);
                    kv = *reinterpret_cast<const uint4*>(p.k + gi * p.ldk + head * HD + c * 8);
                    vv = *reinterpret_cast<const uint4*>(p.v + gi * p.ldv + head * HD + c * 8);
                }
                *reinterpret_cast<uint4*>(sK + j * LDS_ROW + c * 8) = kv;
                *reinterpret_cast<uint4*>(sV + j * LDS_ROW + c * 8) = vv;
            }
            for (int j = tid; j < KC; j += NT) {
                const int key = kc * KC + j;
                float mv = -INFINITY;
                if (key < p.Sk) mv = p.add_mask ? p.add_mask[static_cast<int64_t>(r) * p.Sk + key] * LOG2E : 0.f;
                sM[j] = mv;
            }
            __syncthreads();
            if (!active) continue;

            // ---- S = Q K^T for 16 rows x 64 keys
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                const __nv_bfloat16* kb = sK + (nt * 8 + g) * LDS_ROW + 2 * t;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kb + ks * 16);
                    const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kb + ks * 16 + 8);
                    mma_bf16_16816(s[nt], qf[ks], b0, b1);
                }
            }
            // ---- scale + mask (log2 domain), chunk row max
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float m0 = sM[nt * 8 + 2 * t], m1 = sM[nt * 8 + 2 * t + 1];
                s[nt][0] = fmaf(s[nt][0], scale_log2, m0);
                s[nt][1] = fmaf(s[nt][1], scale_log2, m1);
                s[nt][2] = fmaf(s[nt][2], scale_log2, m0);
                s[nt][3] = fmaf(s[nt][3], scale_log2, m1);
                mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
            }
            float corr[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
                const float m_new = fmaxf(m_run[h], mx[h]);
                corr[h] = (m_run[h] == -INFINITY) ? 0.f : exp2f(m_run[h] - m_new);
                m_run[h] = m_new;
                l_run[h] *= corr[h];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o_acc[i][0] *= corr[0]; o_acc[i][1] *= corr[0];
                o_acc[i][2] *= corr[1]; o_acc[i][3] *= corr[1];
            }
            // ---- P = exp2(S - m), row sums, pack to bf16 A fragments
            uint32_t pf[4][4];
            float ls[2] = {0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = exp2f(s[nt][0] - m_run[0]);
                const float p1 = exp2f(s[nt][1] - m_run[0]);
                const float p2 = exp2f(s[nt][2] - m_run[1]);
                const float p3 = exp2f(s[nt][3] - m_run[1]);
                ls[0] += p0 + p1;
                ls[1] += p2 + p3;
                const int j = nt >> 1;
                if ((nt & 1) == 0) {
                    pf[j][0] = ptx::pack_bf16x2(p0, p1);
                    pf[j][1] = ptx::pack_bf16x2(p2, p3);
                } else {
                    pf[j][2] = ptx::pack_bf16x2(p0, p1);
                    pf[j][3] = ptx::pack_bf16x2(p2, p3);
                }
            }
            l_run[0] += ls[0];
            l_run[1] += ls[1];
            // ---- O += P V   (V^T fragments through ldmatrix.trans from the row-major [key][dim] tile)
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // 16 keys per step
#pragma unroll
                for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide dim tiles
                    uint32_t vf[4];
                    const __nv_bfloat16* vp = sV + (j * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * LDS_ROW + dp * 16 +
                                              (lane >> 4) * 8;
                    ldmatrix_x4_trans(vf, vp);
                    mma_bf16_16816(o_acc[2 * dp], pf[j], vf[0], vf[1]);
                    mma_bf16_16816(o_acc[2 * dp + 1], pf[j], vf[2], vf[3]);
                }
            }
        }
        if (active) {
            // ---- normalise and store
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
                l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
            }
            const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
            const int i0 = rb * 16 + g, i1 = i0 + 8;
            if (i0 < p.Sq) {
                __nv_bfloat16* op = p.o + split_index(r, i0, p.rows, p.nq_split, p.Sq) * p.ldo + head * HD + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(op + nt * 8) = ptx::pack_bf16x2(o_acc[nt][0] * inv0, o_acc[nt][1] * inv0);
            }
            if (i1 < p.Sq) {
                __nv_bfloat16* op = p.o + split_index(r, i1, p.rows, p.nq_split, p.Sq) * p.ldo + head * HD + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(op + nt * 8) = ptx::pack_bf16x2(o_acc[nt][2] * inv1, o_acc[nt][3] * inv1);
            }
        }
    }
}

}  // namespace

int launch_attention(const AttnArgs& a, cudaStream_t s) {
    MRA_REQUIRE(a.rows > 0 && a.heads > 0 && a.Sq > 0 && a.Sk > 0, "attention with empty dimension");
    MRA_REQUIRE(a.Sq <= 512, "attention supports at most 512 query tokens per row, got %d", a.Sq);
    MRA_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0, "attention strides must be 16-byte rows");
    MRA_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.o) & 3) == 0,
                "attention operands must be 16-byte aligned");
    MRA_REQUIRE(static_cast<int64_t>(a.rows) * a.heads < (1ll << 31), "attention grid too large");
    Params p{reinterpret_cast<const __nv_bfloat16*>(a.q), a.ldq, reinterpret_cast<const __nv_bfloat16*>(a.k), a.ldk,
             reinterpret_cast<const __nv_bfloat16*>(a.v), a.ldv, reinterpret_cast<__nv_bfloat16*>(a.o), a.ldo,
             a.add_mask, a.rows, a.heads, a.Sq, a.Sk, a.nq_split, a.kv_dense};
    const int sq_pad = (a.Sq + 15) & ~15;
    const size_t smem = static_cast<size_t>(sq_pad + 2 * KC) * LDS_ROW * 2 + KC * sizeof(float);
    const unsigned grid = static_cast<unsigned>(a.rows) * a.heads;
    static bool attr_set = false;
    if (!attr_set) {
        MRA_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        MRA_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    if (a.Sq <= 32)
        attention_kernel<2><<<grid, 64, smem, s>>>(p);
    else
        attention_kernel<4><<<grid, 128, smem, s>>>(p);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mra
