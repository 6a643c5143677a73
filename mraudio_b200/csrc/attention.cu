// Fused multi-head attention core for the Q-Former (head_dim 64): O = softmax(Q K^T / 8 + mask) V per (row, head).
//
// The problems are tiny (self: 64x64 keys; cross: 32 queries x 257 keys) and there are thousands of them, so this
// is a flash-style register kernel: one CTA per (row, head); each warp owns 16 query rows; keys are streamed through
// shared memory in chunks of 64 with an online softmax, so the [rows, heads, Sq, Sk] score tensor the reference
// materialises (HF port modeling_instructblip.py:512-536) never exists in HBM.  Tensor-core work uses
// mma.sync.m16n8k16 bf16 (fp32 accumulate): at 1.6 % of the path's FLOPs these tiles are too small for tcgen05's
// 128-row atoms to pay off.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

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
    int64_t hsq, hsk, hsv;   // head strides (elements): 64 = heads side by side in a row, else the head-major layout
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
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();

    // ---- stage Q (all Sq rows of this (row, head)); rows >= Sq are zero
    for (int idx = tid; idx < sq_pad * 8; idx += NT) {
        const int i = idx >> 3, c = idx & 7;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (i < p.Sq) {
            const int64_t gi = split_index(r, i, p.rows, p.nq_split, p.Sq);
            val = *reinterpret_cast<const uint4*>(p.q + gi * p.ldq + head * p.hsq + c * 8);
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
                    kv = *reinterpret_cast<const uint4*>(p.k + gi * p.ldk + head * p.hsk + c * 8);
                    vv = *reinterpret_cast<const uint4*>(p.v + gi * p.ldv + head * p.hsv + c * 8);
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


// =====================================================================================================================
// TMA-pipelined variant (the one the forward uses): same math, but Q / K / V tiles arrive as 128B-swizzled 32 x 64 boxes
// through cp.async.bulk.tensor (4-D tensor maps [head][row][token][64 columns], so tokens past the end of a row are
// zero-filled by the TMA unit and never fetched), K/V chunks of 64 keys flow through an mbarrier ring of `stages` slots
// (one for the single-chunk self-attention: more CTAs per SM), and fragments are read with ldmatrix from the swizzled
// tiles (bank-conflict free).  One CTA per (row, head), one warp per 16 queries.  The token axis may consist of two
// segments ("split" layout: query tokens of all rows first, text tokens after them).  The head axis has its own stride:
// 64 elements when the heads sit side by side in a [tokens, heads * 64] matrix, tokens * 64 in the HEAD-MAJOR layout the
// inference forward lets the QKV / cross-K/V GEMMs write ([head][token][64]: every 32-token box is 4 KiB of
// consecutive bytes and a (row, head) tile one contiguous run, instead of 128-byte pieces at the matrix pitch).
// The output rows leave through the finished Q tile: every warp parks its 16 x 64 results in the swizzled tile and writes
// them back as full 128-byte lines (16 bytes per lane) instead of 4-byte pieces.
struct TmaParams {
    __nv_bfloat16* o; int64_t ldo;
    const float* add_mask;
    int rows, heads, Sq, Sk;
    int q_n0, q_n1;   // query tokens per row in segment 0 / 1
    int k_n0, k_n1;   // key tokens per row in segment 0 / 1
    int nchunks;      // ceil(Sk / 64)
    DropoutParams drop;   // thr8 != 0: training-mode dropout on the attention probabilities (dropout.cuh)
};

constexpr int TMA_STAGES_MAX = 3;
constexpr int BOX_BYTES = 32 * 128;          // 32 tokens x 64 bf16
constexpr int STAGE_BYTES_ATT = 4 * BOX_BYTES;  // K lo, K hi, V lo, V hi

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2,
                                            int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            ptx::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(desc)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// address of 16-byte unit `u` of token row `row` inside a tile made of consecutive 32-row swizzled boxes
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int u) {
    return base + row * 128 + ((u ^ (row & 7)) << 4);
}

// up to two independent attention problems per launch (the video and the audio Q-Former of a lockstep forward)
struct TmaMaps {
    CUtensorMap m[12];   // per problem: Q0, Q1, K0, K1, V0, V1
};
struct TmaParams2 {
    TmaParams p[2];
    int split;           // first block of problem 1 (== grid size when there is one problem)
    int stages;          // slots of the K/V ring in shared memory: min(chunks, TMA_STAGES_MAX)
    int o_stage;         // != 0: O rows leave through the Q tile as full 128-byte lines (needs 16-byte aligned output rows)
};

// resident CTAs per SM the register allocation is sized for (shared memory allows as many for the step's shapes)
#ifndef MRA_ATT_MINB4
#define MRA_ATT_MINB4 4
#endif
template <int NWARPS> struct AttOcc { static constexpr int MINB = NWARPS == 2 ? 7 : (NWARPS == 4 ? MRA_ATT_MINB4 : 1); };

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, AttOcc<NWARPS>::MINB)
attention_tma_kernel(const __grid_constant__ TmaMaps maps, const __grid_constant__ TmaParams2 pp) {
    const int prob = static_cast<int>(blockIdx.x) >= pp.split ? 1 : 0;
    const TmaParams& p = pp.p[prob];
    const int bid = static_cast<int>(blockIdx.x) - (prob ? pp.split : 0);
    const CUtensorMap* tmQ0 = &maps.m[prob * 6 + 0];
    const CUtensorMap* tmQ1 = &maps.m[prob * 6 + 1];
    const CUtensorMap* tmK0 = &maps.m[prob * 6 + 2];
    const CUtensorMap* tmK1 = &maps.m[prob * 6 + 3];
    const CUtensorMap* tmV0 = &maps.m[prob * 6 + 4];
    const CUtensorMap* tmV1 = &maps.m[prob * 6 + 5];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int QBOXES = (NWARPS * 16 + 31) / 32;
    const int nstages = pp.stages;
    uint8_t* sQ = smem;                                         // QBOXES boxes
    uint8_t* sKV = smem + QBOXES * BOX_BYTES;                   // nstages x (K 64x64, V 64x64)
    float* sMask = reinterpret_cast<float*>(sKV + nstages * STAGE_BYTES_ATT);   // [nchunks * 64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + p.nchunks * 64);       // q_full, full[nstages]

    const int head = bid % p.heads;
    const int r = bid / p.heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = NWARPS * 32;

    auto load_kv_chunk = [&](int kc, int stage) {
        uint8_t* dst = sKV + stage * STAGE_BYTES_ATT;
        // boxes that start past the last key are neither fetched nor waited for (their shared memory is never read:
        // the tail chunk's MMAs stop at the last 16-key group that holds a key)
        const int nbox = (p.Sk - kc * 64 + 31) / 32 >= 2 ? 2 : 1;
        ptx::mbar_arrive_expect_tx(&bars[1 + stage], 2 * nbox * BOX_BYTES);
        for (int h = 0; h < nbox; ++h) {
            const int key0 = kc * 64 + h * 32;
            const bool seg1 = p.k_n1 > 0 && key0 >= p.k_n0;
            const int tok = seg1 ? key0 - p.k_n0 : key0;   // past the end of the row -> zero fill
            tma_load_4d(dst + h * BOX_BYTES, seg1 ? tmK1 : tmK0, &bars[1 + stage], 0, tok, r, head);
            tma_load_4d(dst + (2 + h) * BOX_BYTES, seg1 ? tmV1 : tmV0, &bars[1 + stage], 0, tok, r, head);
        }
    };

    if (tid == 0) {
        ptx::prefetch_tensormap(tmQ0);
        ptx::prefetch_tensormap(tmK0);
        ptx::prefetch_tensormap(tmV0);
        for (int i = 0; i < 1 + nstages; ++i) ptx::mbar_init(&bars[i], 1);
        ptx::fence_mbar_init();
    }
    ptx::griddep_wait();               // the producers of q / k / v have completed
    ptx::griddep_launch_dependents();
    if (tid == 0) {
        ptx::mbar_arrive_expect_tx(&bars[0], QBOXES * BOX_BYTES);
#pragma unroll
        for (int b = 0; b < QBOXES; ++b) {
            const int q0 = b * 32;
            const bool seg1 = p.q_n1 > 0 && q0 >= p.q_n0;
            tma_load_4d(sQ + b * BOX_BYTES, seg1 ? tmQ1 : tmQ0, &bars[0], 0, seg1 ? q0 - p.q_n0 : q0, r, head);
        }
        for (int s = 0; s < nstages && s < p.nchunks; ++s) load_kv_chunk(s, s);
    }
    // additive mask (log2 domain); -inf past the last key
    for (int j = tid; j < p.nchunks * 64; j += NT) {
        float mv = -INFINITY;
        if (j < p.Sk) mv = p.add_mask ? p.add_mask[static_cast<int64_t>(r) * p.Sk + j] * LOG2E : 0.f;
        sMask[j] = mv;
    }
    __syncthreads();   // barrier inits + mask visible to everyone

    const float scale_log2 = 0.125f * LOG2E;
    const int qr0 = warp * 16;
    const bool active = qr0 < p.Sq;
    const uint32_t qb = ptx::smem_u32(sQ);
    const int qrow = qr0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    ptx::mbar_wait(&bars[0], 0);
    float o_acc[8][4];
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }

    int stage = 0;
    uint32_t phase = 0;
    for (int kc = 0; kc < p.nchunks; ++kc) {
        ptx::mbar_wait(&bars[1 + stage], phase);
        if (active) {
            const uint32_t kb = ptx::smem_u32(sKV + stage * STAGE_BYTES_ATT);
            const uint32_t vb = kb + 2 * BOX_BYTES;
            const float* mk = sMask + kc * 64;
            const int nvalid = p.Sk - kc * 64;   // keys of this chunk (>= 64 except in the tail chunk)
            // ---- S = Q K^T for 16 rows x 64 keys (Q fragments re-read from the tile per chunk: 16 registers fewer held)
            uint32_t qf[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], tile_addr(qb, qrow, ks * 2 + (lane >> 4)));
            float s[8][4];
#pragma unroll
            for (int np = 0; np < 4; ++np) {    // pairs of 8-key tiles
                s[2 * np][0] = s[2 * np][1] = s[2 * np][2] = s[2 * np][3] = 0.f;
                s[2 * np + 1][0] = s[2 * np + 1][1] = s[2 * np + 1][2] = s[2 * np + 1][3] = 0.f;
                if (np * 16 < nvalid) {         // (warp-uniform) 16-key groups past the last key: the mask makes them -inf
                    const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t kf[4];
                        ldmatrix_x4(kf, tile_addr(kb, key, ks * 2 + ((lane >> 3) & 1)));
                        mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                        mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
                    }
                }
            }
            // ---- scale + mask (log2 domain), chunk row max
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 mm = *reinterpret_cast<const float2*>(mk + nt * 8 + 2 * t);
                s[nt][0] = fmaf(s[nt][0], scale_log2, mm.x);
                s[nt][1] = fmaf(s[nt][1], scale_log2, mm.y);
                s[nt][2] = fmaf(s[nt][2], scale_log2, mm.x);
                s[nt][3] = fmaf(s[nt][3], scale_log2, mm.y);
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
            // training-mode dropout: the softmax normaliser (row sum) is that of the UNdropped probabilities; the values that
            // go into P V are masked and rescaled.  One Philox call covers this thread's 16 values of a query row and chunk.
            const bool dropping = p.drop.thr8 != 0;
            uint4 rb0 = make_uint4(0, 0, 0, 0), rb1 = rb0;
            if (dropping) {
                const uint64_t rh = static_cast<uint64_t>(r) * p.heads + head;
                rb0 = dropout_bytes(p.drop, ((rh * p.Sq + (qr0 + g)) * p.nchunks + kc) * 4 + t);
                rb1 = dropout_bytes(p.drop, ((rh * p.Sq + (qr0 + g + 8)) * p.nchunks + kc) * 4 + t);
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float p0 = exp2f(s[nt][0] - m_run[0]);
                float p1 = exp2f(s[nt][1] - m_run[0]);
                float p2 = exp2f(s[nt][2] - m_run[1]);
                float p3 = exp2f(s[nt][3] - m_run[1]);
                ls[0] += p0 + p1;
                ls[1] += p2 + p3;
                if (dropping) {
                    p0 *= dropout_mult(p.drop, rb0, 2 * nt); p1 *= dropout_mult(p.drop, rb0, 2 * nt + 1);
                    p2 *= dropout_mult(p.drop, rb1, 2 * nt); p3 *= dropout_mult(p.drop, rb1, 2 * nt + 1);
                }
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
                if (j * 16 < nvalid) {             // (warp-uniform) P is exactly zero past the last key
                    const int key = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide dim tiles
                        uint32_t vf[4];
                        ldmatrix_x4_t(vf, tile_addr(vb, key, dp * 2 + (lane >> 4)));
                        mma_bf16_16816(o_acc[2 * dp], pf[j], vf[0], vf[1]);
                        mma_bf16_16816(o_acc[2 * dp + 1], pf[j], vf[2], vf[3]);
                    }
                }
            }
        }
        if (kc + nstages < p.nchunks) {
            __syncthreads();   // every warp is done with this stage
            if (tid == 0) load_kv_chunk(kc + nstages, stage);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1; }
    }
    if (active) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        }
        const float inv[2] = {1.f / l_run[0], 1.f / l_run[1]};
        auto out_row = [&](int i) -> __nv_bfloat16* {
            const int64_t grow = (p.q_n1 > 0 && i >= p.q_n0)
                                     ? static_cast<int64_t>(p.rows) * p.q_n0 + static_cast<int64_t>(r) * p.q_n1 + (i - p.q_n0)
                                     : static_cast<int64_t>(r) * p.q_n0 + i;
            return p.o + grow * p.ldo + head * HD;
        };
        if (pp.o_stage) {
            // this warp's 16 rows of the Q tile are finished (only this warp ever read them): park O there, then write
            // each row back as one 128-byte line (8 lanes x 16 bytes)
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = qr0 + g + h * 8;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    ptx::st_shared_b32(tile_addr(qb, row, nt) + 4 * t,
                                       ptx::pack_bf16x2(o_acc[nt][2 * h] * inv[h], o_acc[nt][2 * h + 1] * inv[h]));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = qr0 + i * 4 + (lane >> 3);
                const uint4 val = ptx::ld_shared_v4(tile_addr(qb, row, lane & 7));
                if (row < p.Sq) *reinterpret_cast<uint4*>(out_row(row) + (lane & 7) * 8) = val;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = qr0 + g + h * 8;
                if (i < p.Sq) {
                    __nv_bfloat16* op = out_row(i) + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        *reinterpret_cast<uint32_t*>(op + nt * 8) =
                            ptx::pack_bf16x2(o_acc[nt][2 * h] * inv[h], o_acc[nt][2 * h + 1] * inv[h]);
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn attn_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &res) == cudaSuccess &&
            res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    });
    return fn;
}

struct Map4Key {
    const void* ptr; int64_t ld, hs; int ntok, rows, heads;
    bool operator==(const Map4Key& o) const {
        return ptr == o.ptr && ld == o.ld && hs == o.hs && ntok == o.ntok && rows == o.rows && heads == o.heads;
    }
};
struct Map4Hash {
    size_t operator()(const Map4Key& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.hs);
        h = h * 1000003u ^ static_cast<size_t>(k.heads * 31 + k.ntok);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        return h;
    }
};

// bf16 [heads][rows][ntok][64] view: token stride ld, row stride ntok * ld, head stride hs (all in elements; hs = 64 when the
// heads sit side by side in one [tokens, heads * 64] matrix); box {64, 32, 1, 1}, 128B swizzle.
int get_map4(const void* ptr, int64_t ld, int64_t hs, int ntok, int rows, int heads, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<Map4Key, CUtensorMap, Map4Hash> cache;
    Map4Key key{ptr, ld, hs, ntok, rows, heads};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = attn_encode_fn();
    MRA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(HD), static_cast<cuuint64_t>(ntok), static_cast<cuuint64_t>(rows),
                          static_cast<cuuint64_t>(heads)};
    cuuint64_t gstride[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * ntok, static_cast<cuuint64_t>(hs) * 2};
    cuuint32_t box[4] = {64, 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention) failed with CUresult %d", (int)r);
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 8192) cache.clear();
        cache[key] = m;
    }
    *out = m;
    return 0;
}

template <int NWARPS>
int launch_tma_variant(const TmaMaps& maps, const TmaParams2& pp, unsigned grid, size_t smem, cudaStream_t s) {
    auto kern = attention_tma_kernel<NWARPS>;
    // same L1 / shared-memory split as the GEMM kernels around it: no SM reconfiguration between launches
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(kern), 160 * 1024, true)) return e;
    MRA_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(NWARPS * 32), smem, s, 1, maps, pp));
    return 0;
}

// fills problem slot `slot`; returns -1 when the shape is outside what the TMA kernel covers
int prepare_tma_problem(const AttnArgs& a, int slot, TmaMaps& maps, TmaParams2& pp) {
    const bool split_q = a.nq_split < a.Sq;                  // queries in two segments
    const bool split_k = !a.kv_dense && a.nq_split < a.Sk;   // keys in two segments
    if (a.Sq > 256 || a.Sk > 4096) return -1;
    if ((split_q || split_k) && a.nq_split % 32 != 0) return -1;
    if (a.ldo % 2 != 0) return -1;
    TmaParams& p = pp.p[slot];
    p.o = reinterpret_cast<__nv_bfloat16*>(a.o); p.ldo = a.ldo; p.add_mask = a.add_mask;
    p.rows = a.rows; p.heads = a.heads; p.Sq = a.Sq; p.Sk = a.Sk;
    p.q_n0 = split_q ? a.nq_split : a.Sq; p.q_n1 = split_q ? a.Sq - a.nq_split : 0;
    p.k_n0 = split_k ? a.nq_split : a.Sk; p.k_n1 = split_k ? a.Sk - a.nq_split : 0;
    p.nchunks = (a.Sk + 63) / 64;
    p.drop = a.drop;
    const int64_t hsq = a.hsq ? a.hsq : HD, hsk = a.hsk ? a.hsk : HD, hsv = a.hsv ? a.hsv : HD;
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(a.q);
    const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(a.k);
    const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(a.v);
    CUtensorMap* m = &maps.m[slot * 6];
    if (int e = get_map4(q, a.ldq, hsq, p.q_n0, a.rows, a.heads, &m[0])) return e;
    m[1] = m[0];
    if (p.q_n1 > 0)
        if (int e = get_map4(q + static_cast<int64_t>(a.rows) * p.q_n0 * a.ldq, a.ldq, hsq, p.q_n1, a.rows, a.heads, &m[1])) return e;
    if (int e = get_map4(k, a.ldk, hsk, p.k_n0, a.rows, a.heads, &m[2])) return e;
    m[3] = m[2];
    if (p.k_n1 > 0)
        if (int e = get_map4(k + static_cast<int64_t>(a.rows) * p.k_n0 * a.ldk, a.ldk, hsk, p.k_n1, a.rows, a.heads, &m[3])) return e;
    if (int e = get_map4(v, a.ldv, hsv, p.k_n0, a.rows, a.heads, &m[4])) return e;
    m[5] = m[4];
    if (p.k_n1 > 0)
        if (int e = get_map4(v + static_cast<int64_t>(a.rows) * p.k_n0 * a.ldv, a.ldv, hsv, p.k_n1, a.rows, a.heads, &m[5])) return e;
    return 0;
}

// tuning switches (A/B runs): MRA_ATT_STAGES = slots of the K/V ring (1..3, default 2), MRA_ATT_OSTAGE=0 = 4-byte output stores
int att_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// one launch for `n` (1 or 2) problems with the same number of query warps; -1 = not covered by the TMA kernel
int try_launch_attention_tma(const AttnArgs* a, int n, cudaStream_t s) {
    static const int max_stages = [] { const int v = att_env("MRA_ATT_STAGES", 2); return v < 1 ? 1 : (v > TMA_STAGES_MAX ? TMA_STAGES_MAX : v); }();
    static const int o_stage_on = att_env("MRA_ATT_OSTAGE", 1);
    TmaMaps maps;
    TmaParams2 pp;
    int nw = 0, nchunks = 0;
    unsigned grid = 0;
    bool o_aligned = true;
    for (int i = 0; i < n; ++i) {
        const int e = prepare_tma_problem(a[i], i, maps, pp);
        if (e != 0) return e;
        const int w = (a[i].Sq + 15) / 16;
        const int wc = w <= 2 ? 2 : (w <= 4 ? 4 : (w <= 8 ? 8 : 16));
        if (i > 0 && wc != nw) return -1;
        nw = wc;
        nchunks = pp.p[i].nchunks > nchunks ? pp.p[i].nchunks : nchunks;
        if (i == 0) pp.split = static_cast<int>(a[i].rows) * a[i].heads;
        grid += static_cast<unsigned>(a[i].rows) * a[i].heads;
        o_aligned = o_aligned && (reinterpret_cast<uintptr_t>(a[i].o) & 15) == 0 && a[i].ldo % 8 == 0;
    }
    if (n == 1) {
        pp.p[1] = pp.p[0];
        for (int i = 0; i < 6; ++i) maps.m[6 + i] = maps.m[i];
        pp.split = static_cast<int>(grid);
    }
    pp.stages = nchunks < max_stages ? nchunks : max_stages;
    pp.o_stage = (o_stage_on && o_aligned) ? 1 : 0;
    const int qboxes = (nw * 16 + 31) / 32;
    const size_t smem = static_cast<size_t>(qboxes) * BOX_BYTES + pp.stages * STAGE_BYTES_ATT + static_cast<size_t>(nchunks) * 64 * 4 +
                        (1 + pp.stages) * 8 + 1024;
    if (nw == 2) return launch_tma_variant<2>(maps, pp, grid, smem, s);
    if (nw == 4) return launch_tma_variant<4>(maps, pp, grid, smem, s);
    if (nw == 8) return launch_tma_variant<8>(maps, pp, grid, smem, s);
    return launch_tma_variant<16>(maps, pp, grid, smem, s);
}

}  // namespace

static bool g_force_generic = getenv("MRA_ATTN_GENERIC") != nullptr;
void set_attention_impl_override(int generic) { g_force_generic = generic != 0; }

int launch_attention(const AttnArgs& a, cudaStream_t s) {
    MRA_REQUIRE(a.rows > 0 && a.heads > 0 && a.Sq > 0 && a.Sk > 0, "attention with empty dimension");
    MRA_REQUIRE(a.Sq <= 512, "attention supports at most 512 query tokens per row, got %d", a.Sq);
    MRA_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0, "attention strides must be 16-byte rows");
    MRA_REQUIRE(a.hsq % 8 == 0 && a.hsk % 8 == 0 && a.hsv % 8 == 0 && a.hsq >= 0 && a.hsk >= 0 && a.hsv >= 0,
                "attention head strides must be multiples of 8 elements");
    MRA_REQUIRE(((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a.o) & 3) == 0,
                "attention operands must be 16-byte aligned");
    MRA_REQUIRE(static_cast<int64_t>(a.rows) * a.heads < (1ll << 31), "attention grid too large");
    if (!g_force_generic) {
        const int e = try_launch_attention_tma(&a, 1, s);
        if (e >= 0) return e;
    }
    MRA_REQUIRE(a.drop.thr8 == 0, "attention dropout needs a shape the TMA kernel covers (Sq <= 256, Sk <= 4096, text length a "
                                  "multiple of 32): Sq=%d Sk=%d split=%d", a.Sq, a.Sk, a.nq_split);
    Params p{reinterpret_cast<const __nv_bfloat16*>(a.q), a.ldq, reinterpret_cast<const __nv_bfloat16*>(a.k), a.ldk,
             reinterpret_cast<const __nv_bfloat16*>(a.v), a.ldv, reinterpret_cast<__nv_bfloat16*>(a.o), a.ldo,
             a.add_mask, a.rows, a.heads, a.Sq, a.Sk, a.nq_split, a.kv_dense,
             a.hsq ? a.hsq : HD, a.hsk ? a.hsk : HD, a.hsv ? a.hsv : HD};
    const int sq_pad = (a.Sq + 15) & ~15;
    const size_t smem = static_cast<size_t>(sq_pad + 2 * KC) * LDS_ROW * 2 + KC * sizeof(float);
    const unsigned grid = static_cast<unsigned>(a.rows) * a.heads;
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(attention_kernel<2>), 96 * 1024)) return e;
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(attention_kernel<4>), 96 * 1024)) return e;
    if (a.Sq <= 32)
        MRA_CHECK_CUDA(launch_pdl(attention_kernel<2>, dim3(grid), dim3(64), smem, s, 1, p));
    else
        MRA_CHECK_CUDA(launch_pdl(attention_kernel<4>, dim3(grid), dim3(128), smem, s, 1, p));
    return 0;
}

// Two problems (both Q-Formers of a lockstep forward) in ONE launch when the TMA kernel covers both with the same warp
// count; otherwise one launch each.  *launches receives the number of kernels enqueued.
int launch_attention_pair(const AttnArgs* a, int n, cudaStream_t s, int* launches) {
    if (n == 2 && !g_force_generic) {
        const int e = try_launch_attention_tma(a, 2, s);
        if (e >= 0) {
            if (launches) *launches += 1;
            return e;
        }
    }
    for (int i = 0; i < n; ++i) {
        if (int e = launch_attention(a[i], s)) return e;
        if (launches) *launches += 1;
    }
    return 0;
}

}  // namespace mra
