// Backward-pass kernels of the Q-Former / llm_proj fine-tuning step (the autograd the reference runs through
// utils/trainer.py:129-140 for the trainable parameters; here: Q-Former + projection, encoders and LLM frozen).
// The tensor-core work of the backward (dgrad / wgrad) reuses gemm_tc_kernel; this file holds the HBM-bound pieces:
//   transpose (+ column sums = bias gradients), LayerNorm backward, GELU forward/backward, attention backward,
//   embedding backward, Adam, fp32 -> bf16 casts.
#include "common.h"
#include "ptx.cuh"

namespace mra {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ transpose (+colsum)
// in: bf16 [R, C] (row stride ld_in)  ->  out: bf16 [C, ld_out] with out[c, r] = in[r, c] and zeros for r in [R, ld_out).
// colsum (optional, fp32 [C]) += sum_r in[r, c]   (bias gradient of a Linear whose output gradient is `in`).
__global__ void __launch_bounds__(256)
transpose_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, __nv_bfloat16* __restrict__ out, int64_t ld_out, int R,
                 int C, float* __restrict__ colsum) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8, c = c0 + tx;
        tile[ty + i * 8][tx] = (r < R && c < C) ? __bfloat162float(in[static_cast<int64_t>(r) * ld_in + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, r = r0 + tx;
        if (c < C && r < ld_out) out[static_cast<int64_t>(c) * ld_out + r] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
    }
    if (colsum != nullptr) {
        // warp ty sums column (c0 + lane?) -- each warp reduces 4 columns of the tile over its 32 rows
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cc = ty + i * 8;
            float v = tile[tx][cc];
            v = warp_sum(v);
            if (tx == 0 && c0 + cc < C) atomicAdd(&colsum[c0 + cc], v);
        }
    }
}

// ------------------------------------------------------------------------------------------------ column sums
// colsum[c] += sum_r in[r, c]   (bias gradient of a Linear whose output gradient is `in`); bf16 in, fp32 atomics out.
// Block = 32 column pairs x 8 row lanes over a chunk of 512 rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld, int R, int C, float* __restrict__ colsum) {
    __shared__ float2 red[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + 2 * cx;
    const int r0 = blockIdx.y * 512;
    float2 acc = make_float2(0.f, 0.f);
    if (c < C) {
        const int r1 = min(R, r0 + 512);
        for (int r = r0 + ry; r < r1; r += 8) {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(in + static_cast<int64_t>(r) * ld + c));
            acc.x += v.x;
            acc.y += v.y;
        }
    }
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
        float2 s = red[0][cx];
#pragma unroll
        for (int i = 1; i < 8; ++i) { s.x += red[i][cx].x; s.y += red[i][cx].y; }
        atomicAdd(&colsum[c], s.x);
        atomicAdd(&colsum[c + 1], s.y);
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
constexpr int LNB_WARPS = 8;
constexpr int LNB_MAXV_WIDE = 8;   // columns per lane = 8 * MAXV  (n <= 2048)
constexpr int LNB_MAXV_H = 3;      // n <= 768 (the Q-Former hidden size)

// dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat)),  dyg = dy * gamma, statistics recomputed from `pre`.
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy.   One warp per row, grid-stride over rows.
// DBIAS (narrow rows only, register budget): dbias += sum_rows dx -- the bias gradient of the Linear that produced `pre`
// (its output gradient IS dx), saving a separate column-sum pass over dx.
template <int LNB_MAXV, bool DBIAS>
__global__ void __launch_bounds__(LNB_WARPS * 32)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, const float* __restrict__ gamma,
              float* __restrict__ dx32, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dbias, int rows, int n, float eps, const DropoutParams drop_out,
              const DropoutParams drop_in, int row0) {
    __shared__ float red[LNB_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = n >> 3;
    float ag[LNB_MAXV][8], ab[LNB_MAXV][8], ax[DBIAS ? LNB_MAXV : 1][8];
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ag[i][e] = ab[i][e] = 0.f;
            if (DBIAS) ax[i][e] = 0.f;
        }

    for (int row = blockIdx.x * LNB_WARPS + warp; row < rows; row += gridDim.x * LNB_WARPS) {
        const float* pr = pre + static_cast<int64_t>(row) * n;
        const float* dr = dy + static_cast<int64_t>(row) * n;
        float x[LNB_MAXV][8], d[LNB_MAXV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int vi = i * 32 + lane;
            if (vi < nvec) {
                const float4 a = *reinterpret_cast<const float4*>(pr + vi * 8), b = *reinterpret_cast<const float4*>(pr + vi * 8 + 4);
                x[i][0] = a.x; x[i][1] = a.y; x[i][2] = a.z; x[i][3] = a.w; x[i][4] = b.x; x[i][5] = b.y; x[i][6] = b.z; x[i][7] = b.w;
                const float4 c = *reinterpret_cast<const float4*>(dr + vi * 8), e4 = *reinterpret_cast<const float4*>(dr + vi * 8 + 4);
                d[i][0] = c.x; d[i][1] = c.y; d[i][2] = c.z; d[i][3] = c.w; d[i][4] = e4.x; d[i][5] = e4.y; d[i][6] = e4.z; d[i][7] = e4.w;
                if (drop_in.thr8 != 0) {   // dropout sat on this LayerNorm's output (embeddings): mask the incoming gradient
                    const uint4 rb = dropout_bytes(drop_in, static_cast<uint64_t>(row0 + row) * nvec + vi);
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[i][e] *= dropout_mult(drop_in, rb, e);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) sum += x[i][e];
            }
        }
        const float mean = warp_sum(sum) / n;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i)
            if (i * 32 + lane < nvec)
#pragma unroll
                for (int e = 0; e < 8; ++e) { const float t = x[i][e] - mean; var = fmaf(t, t, var); }
        const float rstd = rsqrtf(warp_sum(var) / n + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int vi = i * 32 + lane;
            if (vi < nvec) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (x[i][e] - mean) * rstd;
                    ag[i][e] = fmaf(d[i][e], xh, ag[i][e]);
                    ab[i][e] += d[i][e];
                    const float dg = d[i][e] * g[e];
                    x[i][e] = xh;
                    d[i][e] = dg;
                    s1 += dg;
                    s2 = fmaf(dg, xh, s2);
                }
            }
        }
        const float c1 = warp_sum(s1) / n, c2 = warp_sum(s2) / n;
#pragma unroll
        for (int i = 0; i < LNB_MAXV; ++i) {
            const int vi = i * 32 + lane;
            if (vi < nvec) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = rstd * (d[i][e] - c1 - x[i][e] * c2);
                if (dx32) {
                    float* p = dx32 + static_cast<int64_t>(row) * n + vi * 8;
                    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
                if (drop_out.thr8 != 0) {
                    // dropout sat between the producing Linear and the residual add: the Linear's output gradient (dx16, and
                    // the bias gradient summed from it) is the masked one; dx32 above -- the residual path -- is not
                    const uint4 rb = dropout_bytes(drop_out, static_cast<uint64_t>(row0 + row) * nvec + vi);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] *= dropout_mult(drop_out, rb, e);
                }
                if (DBIAS) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) ax[i][e] += o[e];
                }
                if (dx16) {
                    uint4 u;
                    u.x = ptx::pack_bf16x2(o[0], o[1]); u.y = ptx::pack_bf16x2(o[2], o[3]);
                    u.z = ptx::pack_bf16x2(o[4], o[5]); u.w = ptx::pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(dx16 + static_cast<int64_t>(row) * n + vi * 8) = u;
                }
            }
        }
    }
    // block reduction of the gamma / beta partial sums: lane's vector slot i covers columns i*256 + lane*8 + e
    if (dgamma == nullptr) return;
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i) {
        if (i * 256 < n) {
#pragma unroll
            for (int half = 0; half < (DBIAS ? 3 : 2); ++half) {
                if (half == 2 && dbias == nullptr) break;
                __syncthreads();
#pragma unroll
                for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = half == 0 ? ag[i][e] : (half == 1 ? ab[i][e] : ax[DBIAS ? i : 0][e]);
                __syncthreads();
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < LNB_WARPS; ++w) v += red[w][threadIdx.x];
                const int col = i * 256 + threadIdx.x;
                if (col < n) atomicAdd((half == 0 ? dgamma : (half == 1 ? dbeta : dbias)) + col, v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ GELU
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    return fmaf(x, pdf, cdf);
}
// y = gelu(z)  (training forward keeps z for the backward);  dz = dy * gelu'(z)
template <bool BWD>
__global__ void __launch_bounds__(256)
gelu_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out, int64_t n8) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 zv = reinterpret_cast<const uint4*>(z)[i];
    uint4 dv = make_uint4(0, 0, 0, 0);
    if (BWD) dv = reinterpret_cast<const uint4*>(dy)[i];
    const uint32_t zz[4] = {zv.x, zv.y, zv.z, zv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = ptx::bf16lo(zz[k]), b = ptx::bf16hi(zz[k]);
        if (BWD) o[k] = ptx::pack_bf16x2(ptx::bf16lo(dd[k]) * gelu_grad(a), ptx::bf16hi(dd[k]) * gelu_grad(b));
        else o[k] = ptx::pack_bf16x2(gelu_exact(a), gelu_exact(b));
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------ attention backward
// One CTA per (row, head), head_dim 64.  Recomputes P = softmax(Q K^T / 8 + mask) (fp32), then
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dP o P)),  dQ = dS K / 8,  dK = dS^T Q / 8.
// Q, K, V, dO tiles live in shared memory as bf16 (row stride 66 elements: conflict-free for "lane = row" and
// "lane = column pair" access), P and dS as fp32 [Sq][Sk].  Token addressing as in the forward (split / dense).
struct AttnBwdParams {
    const __nv_bfloat16* q; int64_t ldq;
    const __nv_bfloat16* k; int64_t ldk;
    const __nv_bfloat16* v; int64_t ldv;
    const __nv_bfloat16* d_o; int64_t ldo;
    __nv_bfloat16* dq; int64_t lddq;
    __nv_bfloat16* dk; int64_t lddk;
    __nv_bfloat16* dv; int64_t lddv;
    const float* add_mask;
    int rows, heads, Sq, Sk, nq_split, kv_dense;
};
constexpr int AB_LD = 66;

__device__ __forceinline__ int64_t tok_index(int r, int i, int rows, int nsplit, int S, bool dense) {
    if (dense) return static_cast<int64_t>(r) * S + i;
    return i < nsplit ? static_cast<int64_t>(r) * nsplit + i
                      : static_cast<int64_t>(rows) * nsplit + static_cast<int64_t>(r) * (S - nsplit) + (i - nsplit);
}

__global__ void __launch_bounds__(256)
attn_bwd_kernel(const AttnBwdParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [Sq][66]
    __nv_bfloat16* sDO = sQ + p.Sq * AB_LD;                           // [Sq][66]
    __nv_bfloat16* sK = sDO + p.Sq * AB_LD;                           // [Sk][66]
    __nv_bfloat16* sV = sK + p.Sk * AB_LD;                            // [Sk][66]
    float* sP = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sV + p.Sk * AB_LD) + 15) & ~uintptr_t(15));
    float* sDS = sP + static_cast<size_t>(p.Sq) * p.Sk;               // [Sq][Sk]
    const int head = blockIdx.x % p.heads, r = blockIdx.x / p.heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool q_dense = p.nq_split >= p.Sq;

    for (int idx = tid; idx < p.Sq * 32; idx += 256) {
        const int i = idx >> 5, c = idx & 31;
        const int64_t gi = tok_index(r, i, p.rows, p.nq_split, p.Sq, q_dense);
        *reinterpret_cast<uint32_t*>(sQ + i * AB_LD + 2 * c) = *reinterpret_cast<const uint32_t*>(p.q + gi * p.ldq + head * 64 + 2 * c);
        *reinterpret_cast<uint32_t*>(sDO + i * AB_LD + 2 * c) = *reinterpret_cast<const uint32_t*>(p.d_o + gi * p.ldo + head * 64 + 2 * c);
    }
    for (int idx = tid; idx < p.Sk * 32; idx += 256) {
        const int j = idx >> 5, c = idx & 31;
        const int64_t gj = tok_index(r, j, p.rows, p.nq_split, p.Sk, p.kv_dense != 0);
        *reinterpret_cast<uint32_t*>(sK + j * AB_LD + 2 * c) = *reinterpret_cast<const uint32_t*>(p.k + gj * p.ldk + head * 64 + 2 * c);
        *reinterpret_cast<uint32_t*>(sV + j * AB_LD + 2 * c) = *reinterpret_cast<const uint32_t*>(p.v + gj * p.ldv + head * 64 + 2 * c);
    }
    __syncthreads();

    // ---- phase 1: one warp per query row: P row, dS row, dQ row
    for (int i = warp; i < p.Sq; i += 8) {
        float* Pi = sP + static_cast<size_t>(i) * p.Sk;
        float* dSi = sDS + static_cast<size_t>(i) * p.Sk;
        float mx = -INFINITY;
        for (int j = lane; j < p.Sk; j += 32) {
            float s = 0.f, dp = 0.f;
#pragma unroll 8
            for (int d = 0; d < 64; d += 2) {
                const float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sQ + i * AB_LD + d));
                const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sK + j * AB_LD + d));
                const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sDO + i * AB_LD + d));
                const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + j * AB_LD + d));
                s = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, s));
                dp = fmaf(ov.x, vv.x, fmaf(ov.y, vv.y, dp));
            }
            s = s * 0.125f + (p.add_mask ? p.add_mask[static_cast<int64_t>(r) * p.Sk + j] : 0.f);
            Pi[j] = s;
            dSi[j] = dp;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < p.Sk; j += 32) {
            const float e = __expf(Pi[j] - mx);
            Pi[j] = e;
            sum += e;
        }
        const float inv = 1.f / warp_sum(sum);
        float dsum = 0.f;
        for (int j = lane; j < p.Sk; j += 32) {
            const float pj = Pi[j] * inv;
            Pi[j] = pj;
            dsum = fmaf(pj, dSi[j], dsum);
        }
        dsum = warp_sum(dsum);
        for (int j = lane; j < p.Sk; j += 32) dSi[j] = Pi[j] * (dSi[j] - dsum);
        __syncwarp();
        // dQ_i[d] = sum_j dS_ij K_j[d] / 8 : lane owns dims 2*lane, 2*lane+1
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < p.Sk; ++j) {
            const float ds = dSi[j];
            const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sK + j * AB_LD + 2 * lane));
            a0 = fmaf(ds, kv.x, a0);
            a1 = fmaf(ds, kv.y, a1);
        }
        const int64_t gi = tok_index(r, i, p.rows, p.nq_split, p.Sq, q_dense);
        *reinterpret_cast<uint32_t*>(p.dq + gi * p.lddq + head * 64 + 2 * lane) = ptx::pack_bf16x2(a0 * 0.125f, a1 * 0.125f);
    }
    __syncthreads();
    // ---- phase 2: one warp per key: dK_j = sum_i dS_ij Q_i / 8, dV_j = sum_i P_ij dO_i
    for (int j = warp; j < p.Sk; j += 8) {
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int i = 0; i < p.Sq; ++i) {
            const float ds = sDS[static_cast<size_t>(i) * p.Sk + j], pj = sP[static_cast<size_t>(i) * p.Sk + j];
            const float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sQ + i * AB_LD + 2 * lane));
            const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sDO + i * AB_LD + 2 * lane));
            k0 = fmaf(ds, qv.x, k0); k1 = fmaf(ds, qv.y, k1);
            v0 = fmaf(pj, ov.x, v0); v1 = fmaf(pj, ov.y, v1);
        }
        const int64_t gj = tok_index(r, j, p.rows, p.nq_split, p.Sk, p.kv_dense != 0);
        *reinterpret_cast<uint32_t*>(p.dk + gj * p.lddk + head * 64 + 2 * lane) = ptx::pack_bf16x2(k0 * 0.125f, k1 * 0.125f);
        *reinterpret_cast<uint32_t*>(p.dv + gj * p.lddv + head * 64 + 2 * lane) = ptx::pack_bf16x2(v0, v1);
    }
}

// ------------------------------------------------------------------------------------------------ attention backward (tensor cores)
// Same contract as attn_bwd_kernel, every product on mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the tiles of one
// (row, head) -- 32..160 queries x 64..272 keys x 64 dims -- are far below a tcgen05 128-row atom, as in the forward.
// One CTA (8 warps) per (row, head); Q, dO, K, V staged once in shared memory (144-byte rows: conflict-free fragment loads
// and ldmatrix), then
//   phase 0  row statistics of S = Q K^T / 8 + mask (online max / sum in the log2 domain, split over key slices when
//            there are fewer 16-query tiles than warps); delta_i = dO_i . O_i from the saved forward output
//   phase 1  per (16 queries x 32 keys) item: S and dP = dO V^T tiles -> P = exp2(S - m) / l and dS = P o (dP - delta)
//            written to shared memory as bf16 [query][key]
//   phase 2  dQ = dS K / 8 (A fragments straight from dS, K^T fragments by ldmatrix.trans), dK = dS^T Q / 8 and
//            dV = P^T dO (A fragments by ldmatrix.trans of the [query][key] tiles): one (16 x 64) output tile per item.
struct AttnBwdTcParams {
    AttnBwdParams b;
    const __nv_bfloat16* o; int64_t ldof;   // forward output (context) of the same attention
    float* db_q; float* db_k; float* db_v;  // optional fused bias gradients of the q / k / v projections: [heads * 64] column sums
    DropoutParams drop;                     // the forward's attention-probability dropout (mask regenerated here)
    int kc;                                 // keys per resident chunk: the padded key count, or a multiple of 64 below it
};
constexpr int ABT_LD = 72;
constexpr int ABT_WARPS = 8;
constexpr float ABT_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void abt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void abt_ldsm_t(uint32_t (&r)[4], const void* smem_row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(ptx::smem_u32(smem_row_ptr)));
}
// A fragments (16 x 64, reduction index contiguous) of rows [r0, r0 + 16) of a staged [rows][ABT_LD] tile
__device__ __forceinline__ void abt_load_a(uint32_t (&f)[4][4], const __nv_bfloat16* tile, int r0, int g, int t) {
    const __nv_bfloat16* b = tile + (r0 + g) * ABT_LD + 2 * t;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        f[ks][0] = *reinterpret_cast<const uint32_t*>(b + ks * 16);
        f[ks][1] = *reinterpret_cast<const uint32_t*>(b + 8 * ABT_LD + ks * 16);
        f[ks][2] = *reinterpret_cast<const uint32_t*>(b + ks * 16 + 8);
        f[ks][3] = *reinterpret_cast<const uint32_t*>(b + 8 * ABT_LD + ks * 16 + 8);
    }
}
// merge two (max, sum) pairs of an online softmax (log2 domain); -inf max = nothing seen yet
__device__ __forceinline__ void abt_merge(float& m, float& l, float m2, float l2) {
    const float mn = fmaxf(m, m2);
    if (mn == -INFINITY) return;
    l = l * exp2f(m - mn) + l2 * exp2f(m2 - mn);
    m = mn;
}

// column sums of a warp's 16 x 64 accumulator tile (rows beyond the valid range are exactly zero) -> shared memory
__device__ __forceinline__ void abt_colsum(const float (&acc)[8][4], float scale, float* scol, int lane) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        float c0 = acc[nt][0] + acc[nt][2], c1 = acc[nt][1] + acc[nt][3];
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if (lane < 4) {
            atomicAdd(scol + nt * 8 + 2 * lane, c0 * scale);
            atomicAdd(scol + nt * 8 + 2 * lane + 1, c1 * scale);
        }
    }
}

__global__ void __launch_bounds__(ABT_WARPS * 32, 3)
attn_bwd_tc_kernel(const AttnBwdTcParams pp) {
    const AttnBwdParams& p = pp.b;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // Keys are processed in chunks of KC (a multiple of 64 when there is more than one chunk; the whole padded key range
    // otherwise): only one chunk of K / V / P / dS is resident, so the cross-attention form (257 keys: chunks of 128) takes
    // 74 KiB instead of 125 KiB of shared memory -- three CTAs per SM instead of one -- and key counts beyond one CTA's
    // shared memory (1024 keys of the Video-LLaMA-style Q-Former) stay on the tensor cores.
    const int Sqp = (p.Sq + 15) & ~15, Skp = (p.Sk + 15) & ~15, KC = pp.kc, NC = (Skp + KC - 1) / KC, ldp = KC + 8;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);     // [Sqp][72]
    __nv_bfloat16* sDO = sQ + Sqp * ABT_LD;                             // [Sqp][72]
    __nv_bfloat16* sK = sDO + Sqp * ABT_LD;                             // [KC][72]   current key chunk
    __nv_bfloat16* sV = sK + KC * ABT_LD;                               // [KC][72]
    __nv_bfloat16* sP = sV + KC * ABT_LD;                               // [Sqp][ldp]
    __nv_bfloat16* sDS = sP + Sqp * ldp;                                // [Sqp][ldp]
    float* sMask = reinterpret_cast<float*>(sDS + Sqp * ldp);           // [Skp]   additive mask * log2e, -inf on padding keys
    float* sDelta = sMask + Skp;                                        // [Sqp]
    float* sMx = sDelta + Sqp;                                          // [Sqp]   row max (log2 domain)
    float* sIl = sMx + Sqp;                                             // [Sqp]   1 / row sum
    float* sPM = sIl + Sqp;                                             // [ABT_WARPS][Sqp] partial max (running over the chunks)
    float* sPL = sPM + ABT_WARPS * Sqp;                                 // [ABT_WARPS][Sqp] partial sum
    float* sCol = sPL + ABT_WARPS * Sqp;                                // [3][64] column sums of dQ / dK / dV (bias gradients)
    float* sDQ = sCol + 192;                                            // [Sqp][64] dQ accumulated over the chunks (NC > 1 only)
    const int head = blockIdx.x % p.heads, r = blockIdx.x / p.heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int NT = ABT_WARPS * 32;
    const bool q_dense = p.nq_split >= p.Sq;

    if (tid < 192) sCol[tid] = 0.f;
    for (int i = tid; i < ABT_WARPS * Sqp; i += NT) { sPM[i] = -INFINITY; sPL[i] = 0.f; }
    __syncthreads();
    // ---- stage Q, dO (+ delta = dO . O); rows past the end are zero
    for (int idx = tid; idx < Sqp * 8; idx += NT) {
        const int i = idx >> 3, c = idx & 7;
        uint4 qv = make_uint4(0, 0, 0, 0), dv = qv, ov = qv;
        if (i < p.Sq) {
            const int64_t gi = tok_index(r, i, p.rows, p.nq_split, p.Sq, q_dense);
            qv = *reinterpret_cast<const uint4*>(p.q + gi * p.ldq + head * 64 + c * 8);
            dv = *reinterpret_cast<const uint4*>(p.d_o + gi * p.ldo + head * 64 + c * 8);
            ov = *reinterpret_cast<const uint4*>(pp.o + gi * pp.ldof + head * 64 + c * 8);
        }
        *reinterpret_cast<uint4*>(sQ + i * ABT_LD + c * 8) = qv;
        *reinterpret_cast<uint4*>(sDO + i * ABT_LD + c * 8) = dv;
        const uint32_t dd[4] = {dv.x, dv.y, dv.z, dv.w}, oo[4] = {ov.x, ov.y, ov.z, ov.w};
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) d = fmaf(ptx::bf16lo(dd[k]), ptx::bf16lo(oo[k]), fmaf(ptx::bf16hi(dd[k]), ptx::bf16hi(oo[k]), d));
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (c == 0) sDelta[i] = d;
        if (pp.db_v && pp.drop.thr8 == 0) {
            // bias gradient of the value projection: sum_j dV_j = sum_q (sum_j P_qj) dO_q = sum_q dO_q (softmax rows sum to 1;
            // with dropout the rows of the masked P do not: the column sums of the dV tiles are taken in phase 2 instead)
            float e8[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { e8[2 * k] = ptx::bf16lo(dd[k]); e8[2 * k + 1] = ptx::bf16hi(dd[k]); }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                e8[k] += __shfl_xor_sync(0xffffffffu, e8[k], 8);
                e8[k] += __shfl_xor_sync(0xffffffffu, e8[k], 16);
            }
            if (lane < 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) atomicAdd(sCol + 128 + c * 8 + k, e8[k]);
            }
        }
    }
    for (int j = tid; j < Skp; j += NT)
        sMask[j] = j < p.Sk ? (p.add_mask ? p.add_mask[static_cast<int64_t>(r) * p.Sk + j] * ABT_LOG2E : 0.f) : -INFINITY;
    // keys [k0, k0 + kn) -> sK (and sV): rows past the end of the row's keys are zero
    auto stage_kv = [&](int k0, int kn, bool with_v) {
        for (int idx = tid; idx < kn * 8; idx += NT) {
            const int jl = idx >> 3, c = idx & 7, j = k0 + jl;
            uint4 kv = make_uint4(0, 0, 0, 0), vv = kv;
            if (j < p.Sk) {
                const int64_t gj = tok_index(r, j, p.rows, p.nq_split, p.Sk, p.kv_dense != 0);
                kv = *reinterpret_cast<const uint4*>(p.k + gj * p.ldk + head * 64 + c * 8);
                if (with_v) vv = *reinterpret_cast<const uint4*>(p.v + gj * p.ldv + head * 64 + c * 8);
            }
            *reinterpret_cast<uint4*>(sK + jl * ABT_LD + c * 8) = kv;
            if (with_v) *reinterpret_cast<uint4*>(sV + jl * ABT_LD + c * 8) = vv;
        }
    };

    const int nQT = Sqp >> 4;
    const float scale2 = 0.125f * ABT_LOG2E;
    // ---- pass 1 over the key chunks: row max / sum (running partials per (query tile, key split) in sPM / sPL)
    const int nsplit = nQT >= ABT_WARPS ? 1 : ABT_WARPS / nQT;
    for (int ck = 0; ck < NC; ++ck) {
        const int k0 = ck * KC, kn = min(KC, Skp - k0), nK8 = kn >> 3;
        if (ck > 0) __syncthreads();              // everybody is done with the previous chunk's keys
        stage_kv(k0, kn, NC == 1);                // a single chunk stays resident for pass 2: stage V with it
        __syncthreads();
        for (int item = warp; item < nQT * nsplit; item += ABT_WARPS) {
            const int qt = item / nsplit, sp = item - qt * nsplit;
            uint32_t qf[4][4];
            abt_load_a(qf, sQ, qt * 16, g, t);
            float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
            for (int nt = sp; nt < nK8; nt += nsplit) {
                float s[4] = {0.f, 0.f, 0.f, 0.f};
                const __nv_bfloat16* kb = sK + (nt * 8 + g) * ABT_LD + 2 * t;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    abt_mma(s, qf[ks], *reinterpret_cast<const uint32_t*>(kb + ks * 16), *reinterpret_cast<const uint32_t*>(kb + ks * 16 + 8));
                const float k0m = sMask[k0 + nt * 8 + 2 * t], k1m = sMask[k0 + nt * 8 + 2 * t + 1];
                const float v[4] = {fmaf(s[0], scale2, k0m), fmaf(s[1], scale2, k1m), fmaf(s[2], scale2, k0m), fmaf(s[3], scale2, k1m)};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float mn = fmaxf(m[h], fmaxf(v[2 * h], v[2 * h + 1]));
                    if (mn != -INFINITY) {
                        l[h] = l[h] * exp2f(m[h] - mn) + exp2f(v[2 * h] - mn) + exp2f(v[2 * h + 1] - mn);
                        m[h] = mn;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const float m2 = __shfl_xor_sync(0xffffffffu, m[h], o), l2 = __shfl_xor_sync(0xffffffffu, l[h], o);
                    abt_merge(m[h], l[h], m2, l2);
                }
                if (t == 0) {   // (the same warp owns this (split, row) entry in every chunk)
                    const int e = sp * Sqp + qt * 16 + g + 8 * h;
                    float rm = sPM[e], rl = sPL[e];
                    abt_merge(rm, rl, m[h], l[h]);
                    sPM[e] = rm;
                    sPL[e] = rl;
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < Sqp; i += NT) {
        float m = -INFINITY, l = 0.f;
        for (int sp = 0; sp < nsplit; ++sp) abt_merge(m, l, sPM[sp * Sqp + i], sPL[sp * Sqp + i]);
        sMx[i] = m == -INFINITY ? 0.f : m;
        sIl[i] = l > 0.f ? 1.f / l : 0.f;
    }
    // ---- pass 2 over the key chunks
    const int lrow = ((lane >> 3) & 1) * 8 + (lane & 7), lcol = (lane >> 4) * 8;     // ldmatrix.trans of a [reduction][dim] tile (B)
    const int arow = ((lane >> 4) & 1) * 8 + (lane & 7), acol = ((lane >> 3) & 1) * 8;   // ... of a [query][key] tile (A^T)
    for (int ck = 0; ck < NC; ++ck) {
        const int k0 = ck * KC, kn = min(KC, Skp - k0), nK8 = kn >> 3, nKT = kn >> 4;
        const bool last_chunk = ck == NC - 1;
        __syncthreads();                          // statistics written / previous chunk's phase 2 done
        if (NC > 1) {
            stage_kv(k0, kn, true);
            __syncthreads();
        }
        // ---- phase 1: P and dS tiles (16 queries x 32 keys per item) -> shared memory, bf16
        const int nCH = (nK8 + 3) >> 2;
        for (int item = warp; item < nQT * nCH; item += ABT_WARPS) {
            const int qt = item / nCH, ch = item - qt * nCH;
            uint32_t qf[4][4], df[4][4];
            abt_load_a(qf, sQ, qt * 16, g, t);
            abt_load_a(df, sDO, qt * 16, g, t);
            const int i0 = qt * 16 + g, i1 = i0 + 8;
            const float m0 = sMx[i0], m1 = sMx[i1], il0 = sIl[i0], il1 = sIl[i1], d0 = sDelta[i0], d1 = sDelta[i1];
            const int nt_end = min(ch * 4 + 4, nK8);
            // forward dropout mask of this thread's elements: one Philox call per query row and 64-key chunk of the row (a
            // 32-key item lies inside one: k0 is a multiple of 64); O = (P o M c) V  =>  dV = (P o M c)^T dO,
            // dP = (dO V^T) o M c,  delta = dO . O unchanged
            const bool dropping = pp.drop.thr8 != 0;
            uint4 rb0 = make_uint4(0, 0, 0, 0), rb1 = rb0;
            if (dropping) {
                const uint64_t rh = static_cast<uint64_t>(r) * p.heads + head;
                const int nchunks = (p.Sk + 63) >> 6, kc = (k0 >> 6) + (ch >> 1);
                rb0 = dropout_bytes(pp.drop, ((rh * p.Sq + i0) * nchunks + kc) * 4 + t);
                rb1 = dropout_bytes(pp.drop, ((rh * p.Sq + i1) * nchunks + kc) * 4 + t);
            }
            for (int nt = ch * 4; nt < nt_end; ++nt) {
                float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                const __nv_bfloat16* kb = sK + (nt * 8 + g) * ABT_LD + 2 * t;
                const __nv_bfloat16* vb = sV + (nt * 8 + g) * ABT_LD + 2 * t;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    abt_mma(s, qf[ks], *reinterpret_cast<const uint32_t*>(kb + ks * 16), *reinterpret_cast<const uint32_t*>(kb + ks * 16 + 8));
                    abt_mma(dp, df[ks], *reinterpret_cast<const uint32_t*>(vb + ks * 16), *reinterpret_cast<const uint32_t*>(vb + ks * 16 + 8));
                }
                const float k0m = sMask[k0 + nt * 8 + 2 * t], k1m = sMask[k0 + nt * 8 + 2 * t + 1];
                const float p0 = exp2f(fmaf(s[0], scale2, k0m) - m0) * il0, p1 = exp2f(fmaf(s[1], scale2, k1m) - m0) * il0;
                const float p2 = exp2f(fmaf(s[2], scale2, k0m) - m1) * il1, p3 = exp2f(fmaf(s[3], scale2, k1m) - m1) * il1;
                const int c = nt * 8 + 2 * t;
                float d0m = 1.f, d1m = 1.f, d2m = 1.f, d3m = 1.f;
                if (dropping) {
                    const int nl = nt & 7;   // 8-key tile inside the 64-key chunk (dynamic index: select the byte at run time)
                    const uint32_t w0 = nl < 2 ? rb0.x : (nl < 4 ? rb0.y : (nl < 6 ? rb0.z : rb0.w));
                    const uint32_t w1 = nl < 2 ? rb1.x : (nl < 4 ? rb1.y : (nl < 6 ? rb1.z : rb1.w));
                    const int sh = (nl & 1) * 16;
                    d0m = ((w0 >> sh) & 0xFFu) >= pp.drop.thr8 ? pp.drop.scale : 0.f;
                    d1m = ((w0 >> (sh + 8)) & 0xFFu) >= pp.drop.thr8 ? pp.drop.scale : 0.f;
                    d2m = ((w1 >> sh) & 0xFFu) >= pp.drop.thr8 ? pp.drop.scale : 0.f;
                    d3m = ((w1 >> (sh + 8)) & 0xFFu) >= pp.drop.thr8 ? pp.drop.scale : 0.f;
                }
                *reinterpret_cast<uint32_t*>(sP + i0 * ldp + c) = ptx::pack_bf16x2(p0 * d0m, p1 * d1m);
                *reinterpret_cast<uint32_t*>(sP + i1 * ldp + c) = ptx::pack_bf16x2(p2 * d2m, p3 * d3m);
                *reinterpret_cast<uint32_t*>(sDS + i0 * ldp + c) = ptx::pack_bf16x2(p0 * (dp[0] * d0m - d0), p1 * (dp[1] * d1m - d0));
                *reinterpret_cast<uint32_t*>(sDS + i1 * ldp + c) = ptx::pack_bf16x2(p2 * (dp[2] * d2m - d1), p3 * (dp[3] * d3m - d1));
            }
        }
        __syncthreads();
        // ---- phase 2: one 16 x 64 output tile per item: dQ tiles (accumulated over the chunks), then this chunk's dK and dV tiles
        for (int item = warp; item < nQT + 2 * nKT; item += ABT_WARPS) {
            float acc[8][4];
            if (item < nQT) {
                const int i0 = item * 16 + g, i1 = i0 + 8;
                float* dq0 = sDQ + i0 * 64 + 2 * t;
                float* dq1 = sDQ + i1 * 64 + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    if (ck > 0) {
                        const float2 a0 = *reinterpret_cast<const float2*>(dq0 + nt * 8), a1 = *reinterpret_cast<const float2*>(dq1 + nt * 8);
                        acc[nt][0] = a0.x; acc[nt][1] = a0.y; acc[nt][2] = a1.x; acc[nt][3] = a1.y;
                    } else {
                        acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
                    }
                }
                for (int kt = 0; kt < nKT; ++kt) {
                    uint32_t a[4];
                    const __nv_bfloat16* ab = sDS + i0 * ldp + kt * 16 + 2 * t;
                    a[0] = *reinterpret_cast<const uint32_t*>(ab);
                    a[1] = *reinterpret_cast<const uint32_t*>(ab + 8 * ldp);
                    a[2] = *reinterpret_cast<const uint32_t*>(ab + 8);
                    a[3] = *reinterpret_cast<const uint32_t*>(ab + 8 * ldp + 8);
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t bf[4];
                        abt_ldsm_t(bf, sK + (kt * 16 + lrow) * ABT_LD + dp * 16 + lcol);
                        abt_mma(acc[2 * dp], a, bf[0], bf[1]);
                        abt_mma(acc[2 * dp + 1], a, bf[2], bf[3]);
                    }
                }
                if (!last_chunk) {
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        *reinterpret_cast<float2*>(dq0 + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
                        *reinterpret_cast<float2*>(dq1 + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
                    }
                    continue;
                }
                if (pp.db_q) abt_colsum(acc, 0.125f, sCol, lane);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = h ? i1 : i0;
                    if (i < p.Sq) {
                        __nv_bfloat16* dst = p.dq + tok_index(r, i, p.rows, p.nq_split, p.Sq, q_dense) * p.lddq + head * 64 + 2 * t;
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt)
                            *reinterpret_cast<uint32_t*>(dst + nt * 8) = ptx::pack_bf16x2(acc[nt][2 * h] * 0.125f, acc[nt][2 * h + 1] * 0.125f);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                const bool is_dk = item < nQT + nKT;
                const int kt = item - nQT - (is_dk ? 0 : nKT);
                const __nv_bfloat16* lhs = is_dk ? sDS : sP;     // [query][key], read transposed
                const __nv_bfloat16* rhs = is_dk ? sQ : sDO;     // [query][dim]
                for (int qt = 0; qt < nQT; ++qt) {
                    uint32_t a[4];
                    abt_ldsm_t(a, lhs + (qt * 16 + arow) * ldp + kt * 16 + acol);
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {
                        uint32_t bf[4];
                        abt_ldsm_t(bf, rhs + (qt * 16 + lrow) * ABT_LD + dp * 16 + lcol);
                        abt_mma(acc[2 * dp], a, bf[0], bf[1]);
                        abt_mma(acc[2 * dp + 1], a, bf[2], bf[3]);
                    }
                }
                if (!is_dk && pp.db_v && pp.drop.thr8 != 0) abt_colsum(acc, 1.0f, sCol + 128, lane);   // (see the staging loop)
                const float sc = is_dk ? 0.125f : 1.0f;
                __nv_bfloat16* out = is_dk ? p.dk : p.dv;
                const int64_t ld = is_dk ? p.lddk : p.lddv;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = k0 + kt * 16 + g + 8 * h;
                    if (j < p.Sk) {
                        __nv_bfloat16* dst = out + tok_index(r, j, p.rows, p.nq_split, p.Sk, p.kv_dense != 0) * ld + head * 64 + 2 * t;
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt)
                            *reinterpret_cast<uint32_t*>(dst + nt * 8) = ptx::pack_bf16x2(acc[nt][2 * h] * sc, acc[nt][2 * h + 1] * sc);
                    }
                }
            }
        }
    }
    // ---- bias gradients: this (row, head)'s column sums -> global (one atomic per column and CTA)
    if (pp.db_q || pp.db_k || pp.db_v) {
        __syncthreads();
        // (the key-projection bias gradient is exactly zero: sum_j dS_qj = 0 for a softmax; db_k is left untouched)
        if (tid < 192 && (tid < 64 || tid >= 128)) {
            float* dst = tid < 64 ? pp.db_q : pp.db_v;
            if (dst) atomicAdd(dst + head * 64 + (tid & 63), sCol[tid]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ embedding backward
// d_emb: fp32 [rows*Nq + rows*T, H] (split layout).  query rows -> d_query[q_rows, Nq, H]; text rows -> scatter-add into
// the word / position embedding gradients.
__global__ void __launch_bounds__(256)
embed_bwd_query_kernel(const float* __restrict__ d_emb, float* __restrict__ d_query, int q_rows, int rows, int Nq, int H) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over Nq*H (q_rows==1) or rows*Nq*H
    const int64_t per = static_cast<int64_t>(Nq) * H;
    if (q_rows == 1) {
        if (idx >= per) return;
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += d_emb[static_cast<int64_t>(r) * per + idx];
        d_query[idx] += s;
    } else {
        if (idx >= per * rows) return;
        d_query[idx] += d_emb[idx];
    }
}
__global__ void __launch_bounds__(256)
embed_bwd_text_kernel(const float* __restrict__ d_emb_text, const int32_t* __restrict__ ids, float* __restrict__ d_word,
                      float* __restrict__ d_pos, int rows, int T, int H, int vocab) {
    const int tok = blockIdx.x;   // r*T + j
    const int j = tok % T;
    int id = ids[tok];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        const float g = d_emb_text[static_cast<int64_t>(tok) * H + c];
        atomicAdd(&d_word[static_cast<int64_t>(id) * H + c], g);
        atomicAdd(&d_pos[static_cast<int64_t>(j) * H + c], g);
    }
}

// ------------------------------------------------------------------------------------------------ optimizer / casts
// torch.optim.Adam semantics (utils/trainer.py:65: Adam(lr) with default betas / eps, L2 weight decay added to the grad).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 x = *reinterpret_cast<const float4*>(in + i);
        uint2 o;
        o.x = ptx::pack_bf16x2(x.x, x.y);
        o.y = ptx::pack_bf16x2(x.z, x.w);
        *reinterpret_cast<uint2*>(out + i) = o;
    } else {
        for (int64_t k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
    }
}

}  // namespace

int launch_transpose(const void* in, int64_t ld_in, void* out, int64_t ld_out, int R, int C, float* colsum, cudaStream_t s) {
    MRA_REQUIRE(R > 0 && C > 0 && ld_out >= R, "transpose: bad shape R=%d C=%d ld_out=%lld", R, C, (long long)ld_out);
    dim3 grid((C + 31) / 32, static_cast<unsigned>((ld_out + 31) / 32));
    transpose_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld_in, reinterpret_cast<__nv_bfloat16*>(out),
                                          ld_out, R, C, colsum);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_colsum(const void* in, int64_t ld, int R, int C, float* colsum, cudaStream_t s) {
    MRA_REQUIRE(R > 0 && C > 0 && C % 2 == 0 && ld % 2 == 0, "colsum: bad shape R=%d C=%d", R, C);
    dim3 grid((C + 63) / 64, (R + 511) / 512);
    colsum_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld, R, C, colsum);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_ln_bwd(const float* dy, const float* pre, const float* gamma, float* dx32, void* dx16, float* dgamma, float* dbeta,
                  float* dbias, int rows, int n, float eps, cudaStream_t s, const DropoutParams* drop_out, const DropoutParams* drop_in,
                  int row0) {
    MRA_REQUIRE(rows > 0 && n % 8 == 0 && n <= 32 * LNB_MAXV_WIDE * 8, "layernorm backward width %d unsupported", n);
    const DropoutParams dout = drop_out ? *drop_out : DropoutParams(), din = drop_in ? *drop_in : DropoutParams();
    MRA_REQUIRE(dbias == nullptr || (n <= 32 * LNB_MAXV_H * 8 && dgamma != nullptr),
                "layernorm backward: the fused bias gradient needs width <= %d and dgamma", 32 * LNB_MAXV_H * 8);
    int blocks = (rows + LNB_WARPS - 1) / LNB_WARPS;
    const int cap = sm_count() * 4;
    if (blocks > cap) blocks = cap;
    __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dx16);
    if (n <= 32 * LNB_MAXV_H * 8)
        ln_bwd_kernel<LNB_MAXV_H, true><<<blocks, LNB_WARPS * 32, 0, s>>>(dy, pre, gamma, dx32, d16, dgamma, dbeta, dbias, rows, n, eps,
                                                                          dout, din, row0);
    else
        ln_bwd_kernel<LNB_MAXV_WIDE, false><<<blocks, LNB_WARPS * 32, 0, s>>>(dy, pre, gamma, dx32, d16, dgamma, dbeta, nullptr, rows, n,
                                                                              eps, dout, din, row0);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// dz = dy * gelu'(z) over [rows, cols] with the bias gradient of the producing Linear fused: colsum[c] += sum_r dz[r, c].
// Block = 128 threads x 8 columns over a chunk of GB_ROWS rows.
constexpr int GB_ROWS = 64;
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ colsum, int rows, int cols) {
    __shared__ float red[8][32][9];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + cx) * 8;
    const int r0 = blockIdx.y * GB_ROWS, r1 = min(rows, r0 + GB_ROWS);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < cols) {
#pragma unroll 4
        for (int r = r0 + ry; r < r1; r += 8) {
            const int64_t off = static_cast<int64_t>(r) * cols + c;
            const uint4 zv = *reinterpret_cast<const uint4*>(z + off), dv = *reinterpret_cast<const uint4*>(dy + off);
            const uint32_t zz[4] = {zv.x, zv.y, zv.z, zv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k] = ptx::pack_bf16x2(ptx::bf16lo(dd[k]) * gelu_grad(ptx::bf16lo(zz[k])), ptx::bf16hi(dd[k]) * gelu_grad(ptx::bf16hi(zz[k])));
                acc[2 * k] += ptx::bf16lo(o[k]);      // the rounded values, i.e. exactly the column sums of what the wgrad GEMM reads
                acc[2 * k + 1] += ptx::bf16hi(o[k]);
            }
            *reinterpret_cast<uint4*>(out + off) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[ry][cx][e] = acc[e];
    __syncthreads();
    // 256 threads reduce the 8 row lanes of the block's 256 columns
    const int col = threadIdx.x, vcx = col >> 3, ve = col & 7;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][vcx][ve];
    const int gc = blockIdx.x * 256 + col;
    if (gc < cols) atomicAdd(colsum + gc, v);
}

int launch_gelu_bwd_colsum(const void* z, const void* dy, void* dz, float* colsum, int rows, int cols, cudaStream_t s) {
    MRA_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && colsum, "gelu backward: bad shape rows=%d cols=%d", rows, cols);
    dim3 grid((cols + 255) / 256, (rows + GB_ROWS - 1) / GB_ROWS);
    gelu_bwd_colsum_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(z), reinterpret_cast<const __nv_bfloat16*>(dy),
                                                reinterpret_cast<__nv_bfloat16*>(dz), colsum, rows, cols);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_gelu_fwd(const void* z, void* out, int64_t n, cudaStream_t s) {
    MRA_REQUIRE(n % 8 == 0, "gelu: element count must be a multiple of 8");
    const int64_t n8 = n / 8;
    gelu_kernel<false><<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(z), nullptr,
                                                                             reinterpret_cast<__nv_bfloat16*>(out), n8);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}
int launch_gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, cudaStream_t s) {
    MRA_REQUIRE(n % 8 == 0, "gelu: element count must be a multiple of 8");
    const int64_t n8 = n / 8;
    gelu_kernel<true><<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                                                            reinterpret_cast<const __nv_bfloat16*>(dy),
                                                                            reinterpret_cast<__nv_bfloat16*>(dz), n8);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static size_t attn_bwd_tc_smem(int Sq, int Sk, int kc) {
    const size_t Sqp = (Sq + 15) & ~15, Skp = (Sk + 15) & ~15, ldp = kc + 8;
    const size_t dq = Skp > static_cast<size_t>(kc) ? Sqp * 64 : 0;
    return (2 * Sqp + 2 * kc) * ABT_LD * 2 + 2 * Sqp * ldp * 2 + (Skp + 3 * Sqp + 2 * ABT_WARPS * Sqp + 192 + dq) * 4;
}
// Keys per resident chunk: everything when the padded key count is <= 192 (self-attention, short encoders); otherwise
// equal chunks of a multiple of 64 keys (the dropout mask is generated per 64-key chunk), at most 128 -- and fewer if the
// query tile is large -- so that the CTA stays within 220 KiB.  Measured on the cross-attention form (32 queries x 257 keys,
// 768 CTAs): whole key range resident 126 us (125 KiB, one CTA per SM), chunks of 192 keys 103 us (two per SM), 128 keys
// 84 us (74 KiB, three per SM), 64 keys 97 us.  0: no chunk size fits (the fp32 CUDA-core kernel takes over).
static int attn_bwd_tc_chunk(int Sq, int Sk) {
    static const int forced = [] { const char* e = getenv("MRA_ATTN_BWD_KC"); return e ? atoi(e) : 0; }();
    const int Skp = (Sk + 15) & ~15;
    if (Skp <= 192 && attn_bwd_tc_smem(Sq, Sk, Skp) <= 220 * 1024) return Skp;
    for (int cap = forced > 0 ? forced : 128; cap >= 64; cap -= 64) {
        const int nc = (Skp + cap - 1) / cap;
        const int kc = (((Skp + nc - 1) / nc) + 63) & ~63;
        if (kc < Skp && attn_bwd_tc_smem(Sq, Sk, kc) <= 220 * 1024) return kc;
    }
    return 0;
}

int launch_attention_bwd(const AttnBwdArgs& a, cudaStream_t s) {
    MRA_REQUIRE(a.rows > 0 && a.heads > 0 && a.Sq > 0 && a.Sk > 0, "attention backward with empty dimension");
    // tensor-core kernel when the forward output is available and the tiles fit in shared memory; MRA_ATTN_BWD_SIMT=1
    // forces the fp32 CUDA-core kernel (tests compare the two)
    static const bool force_simt = getenv("MRA_ATTN_BWD_SIMT") != nullptr;
    const int kc = attn_bwd_tc_chunk(a.Sq, a.Sk);
    const size_t tc_smem = kc > 0 ? attn_bwd_tc_smem(a.Sq, a.Sk, kc) : 0;
    if (a.o != nullptr && !force_simt && kc > 0) {
        if (int e = ensure_smem_attr(reinterpret_cast<const void*>(attn_bwd_tc_kernel), 220 * 1024)) return e;
        AttnBwdTcParams pp{{reinterpret_cast<const __nv_bfloat16*>(a.q), a.ldq, reinterpret_cast<const __nv_bfloat16*>(a.k), a.ldk,
                            reinterpret_cast<const __nv_bfloat16*>(a.v), a.ldv, reinterpret_cast<const __nv_bfloat16*>(a.d_o), a.ldo,
                            reinterpret_cast<__nv_bfloat16*>(a.dq), a.lddq, reinterpret_cast<__nv_bfloat16*>(a.dk), a.lddk,
                            reinterpret_cast<__nv_bfloat16*>(a.dv), a.lddv, a.add_mask, a.rows, a.heads, a.Sq, a.Sk, a.nq_split, a.kv_dense},
                           reinterpret_cast<const __nv_bfloat16*>(a.o), a.ldof, a.db_q, a.db_k, a.db_v, a.drop, kc};
        attn_bwd_tc_kernel<<<static_cast<unsigned>(a.rows) * a.heads, ABT_WARPS * 32, tc_smem, s>>>(pp);
        MRA_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    MRA_REQUIRE(a.drop.thr8 == 0, "attention backward with dropout needs the tensor-core kernel (forward output given, tiles within "
                                  "220 KiB of shared memory): Sq=%d Sk=%d", a.Sq, a.Sk);
    const size_t smem = static_cast<size_t>(2 * a.Sq + 2 * a.Sk) * AB_LD * 2 + 16 + static_cast<size_t>(a.Sq) * a.Sk * 8;
    MRA_REQUIRE(smem <= 220 * 1024, "attention backward: Sq=%d x Sk=%d needs %zu bytes of shared memory (max 220 KiB)", a.Sq, a.Sk, smem);
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(attn_bwd_kernel), 220 * 1024)) return e;
    AttnBwdParams p{reinterpret_cast<const __nv_bfloat16*>(a.q), a.ldq, reinterpret_cast<const __nv_bfloat16*>(a.k), a.ldk,
                    reinterpret_cast<const __nv_bfloat16*>(a.v), a.ldv, reinterpret_cast<const __nv_bfloat16*>(a.d_o), a.ldo,
                    reinterpret_cast<__nv_bfloat16*>(a.dq), a.lddq, reinterpret_cast<__nv_bfloat16*>(a.dk), a.lddk,
                    reinterpret_cast<__nv_bfloat16*>(a.dv), a.lddv, a.add_mask, a.rows, a.heads, a.Sq, a.Sk, a.nq_split, a.kv_dense};
    attn_bwd_kernel<<<static_cast<unsigned>(a.rows) * a.heads, 256, smem, s>>>(p);
    MRA_CHECK_CUDA(cudaGetLastError());
    // bias gradients by separate column-sum passes on this path
    const int HC = a.heads * 64;
    if (a.db_q) { if (int e = launch_colsum(a.dq, a.lddq, static_cast<int>(a.n_q_tokens), HC, a.db_q, s)) return e; }
    if (a.db_k) { if (int e = launch_colsum(a.dk, a.lddk, static_cast<int>(a.n_k_tokens), HC, a.db_k, s)) return e; }
    if (a.db_v) { if (int e = launch_colsum(a.dv, a.lddv, static_cast<int>(a.n_k_tokens), HC, a.db_v, s)) return e; }
    return 0;
}

int launch_embed_bwd(const float* d_emb, const int32_t* ids, float* d_query, int q_rows, float* d_word, float* d_pos, int rows,
                     int Nq, int T, int H, int vocab, cudaStream_t s) {
    if (d_query) {
        const int64_t n = static_cast<int64_t>(Nq) * H * (q_rows == 1 ? 1 : rows);
        embed_bwd_query_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(d_emb, d_query, q_rows, rows, Nq, H);
        MRA_CHECK_CUDA(cudaGetLastError());
    }
    if (T > 0 && d_word && d_pos) {
        embed_bwd_text_kernel<<<static_cast<unsigned>(rows) * T, 256, 0, s>>>(d_emb + static_cast<int64_t>(rows) * Nq * H, ids, d_word,
                                                                          d_pos, rows, T, H, vocab);
        MRA_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}

// Same update with the two passes that always follow it folded in: the bf16 operand copy of the new parameters (what the
// GEMMs read) and the reset of the gradient accumulator -- 7 instead of 10 fp32 streams over the 186 M parameters.
__global__ void __launch_bounds__(256)
adam_fused_kernel(float4* __restrict__ p, float4* __restrict__ g, const uint2* __restrict__ g16, float4* __restrict__ m,
                  float4* __restrict__ v, uint2* __restrict__ p16, int64_t n4, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float bc1, float bc2_sqrt, float grad_scale, int zero_grad, const float* __restrict__ dyn) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    if (dyn != nullptr) {   // per-step scalars from device memory: the launch can be replayed from a CUDA graph
        lr = __ldg(dyn); bc1 = __ldg(dyn + 1); bc2_sqrt = __ldg(dyn + 2); grad_scale = __ldg(dyn + 3);
    }
    const float4 m4 = m[i], v4 = v[i];
    float4 p4 = p[i];
    float gg[4];
    if (g16 != nullptr) {   // gradients as reduced over the ranks in bf16 (the fp32 accumulator is only reset)
        const uint2 gq = g16[i];
        gg[0] = ptx::bf16lo(gq.x); gg[1] = ptx::bf16hi(gq.x); gg[2] = ptx::bf16lo(gq.y); gg[3] = ptx::bf16hi(gq.y);
    } else {
        const float4 g4 = g[i];
        gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
    }
    const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mo[4], vo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float gi = gg[k] * grad_scale;
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pp[k], gi);
        mo[k] = fmaf(beta1, mm[k], (1.f - beta1) * gi);
        vo[k] = fmaf(beta2, vv[k], (1.f - beta2) * gi * gi);
        const float denom = sqrtf(vo[k]) / bc2_sqrt + eps;
        pp[k] = pp[k] - (lr / bc1) * (mo[k] / denom);
    }
    m[i] = make_float4(mo[0], mo[1], mo[2], mo[3]);
    v[i] = make_float4(vo[0], vo[1], vo[2], vo[3]);
    p[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    if (p16) p16[i] = make_uint2(ptx::pack_bf16x2(pp[0], pp[1]), ptx::pack_bf16x2(pp[2], pp[3]));
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int launch_adam_fused(float* p, float* g, const void* g16, float* m, float* v, void* p16, int64_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_scale, int zero_grad, cudaStream_t s, const float* dyn) {
    MRA_REQUIRE(n > 0 && n % 4 == 0 && (step >= 1 || dyn != nullptr), "fused adam: n must be a positive multiple of 4");
    if (dyn != nullptr) step = 1;   // (lr, bias corrections and gradient scale are read from `dyn` by the kernel)
    MRA_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0 && ((reinterpret_cast<uintptr_t>(p16) | reinterpret_cast<uintptr_t>(g16)) & 7) == 0,
                "fused adam: buffers must be 16-byte aligned");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    const int64_t n4 = n / 4;
    adam_fused_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g), reinterpret_cast<const uint2*>(g16), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(v), reinterpret_cast<uint2*>(p16), n4, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale, zero_grad,
        dyn);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

void adam_hyper(float lr, float beta1, float beta2, int step, float grad_scale, float* out4) {
    out4[0] = lr;
    out4[1] = 1.f - powf(beta1, static_cast<float>(step));
    out4[2] = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    out4[3] = grad_scale;
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, float grad_scale, cudaStream_t s) {
    MRA_REQUIRE(n > 0 && step >= 1, "adam: bad arguments");
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
    adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                                      grad_scale);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_cast_bf16(const float* in, void* out, int64_t n, cudaStream_t s) {
    MRA_REQUIRE(n > 0, "cast: empty");
    const int64_t n4 = (n + 3) / 4;
    cast_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mra
