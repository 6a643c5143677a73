// HBM-bound row kernels of the Q-Former path: LayerNorm (fp32 residual stream in, fp32 + bf16 out), the embedding
// gather + LayerNorm, the modality LayerNorm with the frame fold, additive-mask construction and the final re-layout.
// One warp per row, 16-byte vectorised coalesced accesses, warp-shuffle reductions, two-pass (mean, then centred
// variance) statistics in fp32 -- the same arithmetic as torch.nn.functional.layer_norm.
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"

namespace mra {
namespace {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int MAX_VEC = 8;  // supports row length up to 32 lanes * 8 vec * 8 elems = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x[MAX_VEC][8] holds this lane's elements: vector i covers columns (i*32 + lane)*8 .. +7 (valid when < n).
__device__ __forceinline__ void ln_normalise_store(float (&x)[MAX_VEC][8], int n, int lane, const float* __restrict__ g,
                                                   const float* __restrict__ b, float eps, float* y32,
                                                   __nv_bfloat16* y16, __nv_bfloat16* ylo = nullptr,
                                                   const DropoutParams* drop = nullptr, int64_t drop_row = 0) {
    const int nvec = n >> 3;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i)
        if (i * 32 + lane < nvec)
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += x[i][e];
    const float mean = warp_sum(sum) / n;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i)
        if (i * 32 + lane < nvec)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = x[i][e] - mean;
                var = fmaf(d, d, var);
            }
    const float rstd = rsqrtf(warp_sum(var) / n + eps);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            const int c = vi * 8;
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c)), g1 = __ldg(reinterpret_cast<const float4*>(g + c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
            float y[8];
            y[0] = (x[i][0] - mean) * rstd * g0.x + b0.x;
            y[1] = (x[i][1] - mean) * rstd * g0.y + b0.y;
            y[2] = (x[i][2] - mean) * rstd * g0.z + b0.z;
            y[3] = (x[i][3] - mean) * rstd * g0.w + b0.w;
            y[4] = (x[i][4] - mean) * rstd * g1.x + b1.x;
            y[5] = (x[i][5] - mean) * rstd * g1.y + b1.y;
            y[6] = (x[i][6] - mean) * rstd * g1.z + b1.z;
            y[7] = (x[i][7] - mean) * rstd * g1.w + b1.w;
            if (drop != nullptr) {   // training-mode dropout on the LayerNorm output (embeddings)
                const uint4 rb = dropout_bytes(*drop, static_cast<uint64_t>(drop_row) * nvec + vi);
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] *= dropout_mult(*drop, rb, e);
            }
            if (y32) {
                *reinterpret_cast<float4*>(y32 + c) = make_float4(y[0], y[1], y[2], y[3]);
                *reinterpret_cast<float4*>(y32 + c + 4) = make_float4(y[4], y[5], y[6], y[7]);
            }
            if (y16) {
                uint4 o;
                o.x = ptx::pack_bf16x2(y[0], y[1]);
                o.y = ptx::pack_bf16x2(y[2], y[3]);
                o.z = ptx::pack_bf16x2(y[4], y[5]);
                o.w = ptx::pack_bf16x2(y[6], y[7]);
                *reinterpret_cast<uint4*>(y16 + c) = o;
                if (ylo) {   // split residual stream: lo = bf16(y - hi)
                    uint4 l;
                    l.x = ptx::pack_bf16x2(y[0] - ptx::bf16lo(o.x), y[1] - ptx::bf16hi(o.x));
                    l.y = ptx::pack_bf16x2(y[2] - ptx::bf16lo(o.y), y[3] - ptx::bf16hi(o.y));
                    l.z = ptx::pack_bf16x2(y[4] - ptx::bf16lo(o.z), y[5] - ptx::bf16hi(o.z));
                    l.w = ptx::pack_bf16x2(y[6] - ptx::bf16lo(o.w), y[7] - ptx::bf16hi(o.w));
                    *reinterpret_cast<uint4*>(ylo + c) = l;
                }
            }
        }
    }
}

__device__ __forceinline__ void load8_f32(const float* p, float (&x)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&x)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    x[0] = ptx::bf16lo(v.x); x[1] = ptx::bf16hi(v.x); x[2] = ptx::bf16lo(v.y); x[3] = ptx::bf16hi(v.y);
    x[4] = ptx::bf16lo(v.z); x[5] = ptx::bf16hi(v.z); x[6] = ptx::bf16lo(v.w); x[7] = ptx::bf16hi(v.w);
}
__device__ __forceinline__ void load8_f16(const __half* p, float (&x)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        x[2 * i] = f.x; x[2 * i + 1] = f.y;
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                 float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, int rows, int n, float eps) {
    const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<int64_t>(row) * n;
    float v[MAX_VEC][8];
    const int nvec = n >> 3;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i)
        if (i * 32 + lane < nvec) load8_f32(xr + (i * 32 + lane) * 8, v[i]);
    ln_normalise_store(v, n, lane, g, b, eps, y32 ? y32 + static_cast<int64_t>(row) * n : nullptr,
                       y16 ? y16 + static_cast<int64_t>(row) * n : nullptr);
}

// up to 4 independent row ranges (same width) in one launch: warp w of the grid handles global row w
struct LnGroup {
    LnSegment seg[4];
    int start[5];
    int n;
};
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
layernorm_grouped_kernel(const LnGroup grp, int n, float eps) {
    const int grow = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();
    if (grow >= grp.start[grp.n]) return;
    int g = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < grp.n && grow >= grp.start[i]) g = i;
    const LnSegment& sg = grp.seg[g];
    const int row = grow - grp.start[g];
    const float* xr = sg.x + static_cast<int64_t>(row) * n;
    float v[MAX_VEC][8];
    const int nvec = n >> 3;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i)
        if (i * 32 + lane < nvec) load8_f32(xr + (i * 32 + lane) * 8, v[i]);
    ln_normalise_store(v, n, lane, sg.gamma, sg.beta, eps, sg.y32 ? sg.y32 + static_cast<int64_t>(row) * n : nullptr,
                       sg.y16 ? reinterpret_cast<__nv_bfloat16*>(sg.y16) + static_cast<int64_t>(row) * n : nullptr);
}

// out row (b*F + f) <- in row (frame_major ? f*bs + b : b*F + f), token tok.
template <int DTYPE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
modality_ln_kernel(const void* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                   __nv_bfloat16* __restrict__ out, int bs, int frames, int Nk, int W, int frame_major, float eps) {
    const int64_t token = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t total = static_cast<int64_t>(bs) * frames * Nk;
    if (token >= total) return;
    const int64_t orow = token / Nk;
    const int tok = static_cast<int>(token % Nk);
    const int bidx = static_cast<int>(orow / frames), f = static_cast<int>(orow % frames);
    const int64_t irow = frame_major ? static_cast<int64_t>(f) * bs + bidx : orow;
    const int64_t ioff = (irow * Nk + tok) * W;
    float v[MAX_VEC][8];
    const int nvec = W >> 3;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
            if (DTYPE == 0) load8_f32(reinterpret_cast<const float*>(x) + ioff + vi * 8, v[i]);
            if (DTYPE == 1) load8_bf16(reinterpret_cast<const __nv_bfloat16*>(x) + ioff + vi * 8, v[i]);
            if (DTYPE == 2) load8_f16(reinterpret_cast<const __half*>(x) + ioff + vi * 8, v[i]);
        }
    }
    ln_normalise_store(v, W, lane, g, b, eps, nullptr, out + token * W);
}

// out[b, f, t, :] = bf16(x[b, f, t, :] + pos[f, :])  -- Video-LLaMA-v1 frame position embedding (broadcast over tokens)
template <int DTYPE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
add_frame_pos_kernel(const void* __restrict__ x, const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, int64_t tokens,
                     int frames, int n, int W) {
    const int64_t token = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (token >= tokens) return;
    const int f = static_cast<int>((token / n) % frames);
    const int nvec = W >> 3;
    for (int vi = lane; vi < nvec; vi += 32) {
        float v[8], pe[8];
        const int64_t off = token * W + vi * 8;
        if (DTYPE == 0) load8_f32(reinterpret_cast<const float*>(x) + off, v);
        if (DTYPE == 1) load8_bf16(reinterpret_cast<const __nv_bfloat16*>(x) + off, v);
        if (DTYPE == 2) load8_f16(reinterpret_cast<const __half*>(x) + off, v);
        load8_f32(pos + static_cast<int64_t>(f) * W + vi * 8, pe);
        uint4 o;
        o.x = ptx::pack_bf16x2(v[0] + pe[0], v[1] + pe[1]);
        o.y = ptx::pack_bf16x2(v[2] + pe[2], v[3] + pe[3]);
        o.z = ptx::pack_bf16x2(v[4] + pe[4], v[5] + pe[5]);
        o.w = ptx::pack_bf16x2(v[6] + pe[6], v[7] + pe[7]);
        *reinterpret_cast<uint4*>(out + off) = o;
    }
}

// rows of the split layout: [0, rows*Nq) query tokens (row r, token i), then [rows*Nq, rows*(Nq+T)) text tokens.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
embed_ln_kernel(const float* __restrict__ query_embeds, int q_rows, const int32_t* __restrict__ ids,
                const __nv_bfloat16* __restrict__ word_emb, const __nv_bfloat16* __restrict__ pos_emb,
                const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ y32,
                __nv_bfloat16* __restrict__ y16, __nv_bfloat16* __restrict__ ylo, float* __restrict__ pre_out, int rows, int Nq,
                int T, int H, int vocab, float eps, const DropoutParams drop) {
    const int64_t orow = static_cast<int64_t>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int64_t nquery = static_cast<int64_t>(rows) * Nq;
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();
    if (orow >= nquery + static_cast<int64_t>(rows) * T) return;
    float v[MAX_VEC][8];
    const int nvec = H >> 3;
    if (orow < nquery) {
        const int r = static_cast<int>(orow / Nq), i = static_cast<int>(orow % Nq);
        const float* src = query_embeds + (static_cast<int64_t>(q_rows == 1 ? 0 : r) * Nq + i) * H;
#pragma unroll
        for (int k = 0; k < MAX_VEC; ++k)
            if (k * 32 + lane < nvec) load8_f32(src + (k * 32 + lane) * 8, v[k]);
    } else {
        const int64_t tt = orow - nquery;
        const int r = static_cast<int>(tt / T), j = static_cast<int>(tt % T);
        int id = ids[static_cast<int64_t>(r) * T + j];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const __nv_bfloat16* we = word_emb + static_cast<int64_t>(id) * H;
        const __nv_bfloat16* pe = pos_emb + static_cast<int64_t>(j) * H;
#pragma unroll
        for (int k = 0; k < MAX_VEC; ++k)
            if (k * 32 + lane < nvec) {
                float a[8], c[8];
                load8_bf16(we + (k * 32 + lane) * 8, a);
                load8_bf16(pe + (k * 32 + lane) * 8, c);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k][e] = a[e] + c[e];
            }
    }
    if (pre_out != nullptr) {   // saved for the backward of the embedding LayerNorm
#pragma unroll
        for (int k = 0; k < MAX_VEC; ++k)
            if (k * 32 + lane < nvec) {
                float* pp = pre_out + orow * H + (k * 32 + lane) * 8;
                *reinterpret_cast<float4*>(pp) = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                *reinterpret_cast<float4*>(pp + 4) = make_float4(v[k][4], v[k][5], v[k][6], v[k][7]);
            }
    }
    ln_normalise_store(v, H, lane, g, b, eps, y32 ? y32 + orow * H : nullptr, y16 + orow * H, ylo ? ylo + orow * H : nullptr,
                       drop.thr8 ? &drop : nullptr, orow);
}

__global__ void enc_mask_kernel(const int32_t* __restrict__ enc_mask, float* __restrict__ out, int64_t n) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    out[idx] = (1.0f - static_cast<float>(enc_mask[idx])) * -10000.0f;
}

__global__ void gather_last_hidden_kernel(const float4* __restrict__ split, float4* __restrict__ out, int rows, int Nq, int T,
                                          int H4) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int S = Nq + T;
    const int64_t total = static_cast<int64_t>(rows) * S * H4;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % H4);
    const int64_t tok = idx / H4;
    const int r = static_cast<int>(tok / S), i = static_cast<int>(tok % S);
    const int64_t srow = i < Nq ? static_cast<int64_t>(r) * Nq + i
                                : static_cast<int64_t>(rows) * Nq + static_cast<int64_t>(r) * T + (i - Nq);
    out[idx] = split[srow * H4 + c];
}

// same for the split (hi + lo bf16) residual stream: out = float(hi) + float(lo); 8 elements per thread
__global__ void gather_last_hidden_split_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo, float4* __restrict__ out,
                                                int rows, int Nq, int T, int H8) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int S = Nq + T;
    const int64_t total = static_cast<int64_t>(rows) * S * H8;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % H8);
    const int64_t tok = idx / H8;
    const int r = static_cast<int>(tok / S), i = static_cast<int>(tok % S);
    const int64_t srow = i < Nq ? static_cast<int64_t>(r) * Nq + i
                                : static_cast<int64_t>(rows) * Nq + static_cast<int64_t>(r) * T + (i - Nq);
    const uint4 h = hi[srow * H8 + c], l = lo[srow * H8 + c];
    out[2 * idx] = make_float4(ptx::bf16lo(h.x) + ptx::bf16lo(l.x), ptx::bf16hi(h.x) + ptx::bf16hi(l.x),
                               ptx::bf16lo(h.y) + ptx::bf16lo(l.y), ptx::bf16hi(h.y) + ptx::bf16hi(l.y));
    out[2 * idx + 1] = make_float4(ptx::bf16lo(h.z) + ptx::bf16lo(l.z), ptx::bf16hi(h.z) + ptx::bf16hi(l.z),
                                   ptx::bf16lo(h.w) + ptx::bf16lo(l.w), ptx::bf16hi(h.w) + ptx::bf16hi(l.w));
}

}  // namespace

int launch_layernorm(const float* x, const float* g, const float* b, float* y32, void* y16, int rows, int n, float eps,
                     cudaStream_t s) {
    MRA_REQUIRE(rows > 0 && n > 0 && n % 8 == 0 && n <= 32 * MAX_VEC * 8, "layernorm width %d unsupported (multiple of 8, <= 2048)", n);
    const int blocks = (rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    layernorm_kernel<<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, g, b, y32, reinterpret_cast<__nv_bfloat16*>(y16), rows, n, eps);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_layernorm_grouped(const LnSegment* segs, int nseg, int n, float eps, cudaStream_t s) {
    MRA_REQUIRE(nseg >= 1 && nseg <= 4, "grouped layernorm takes 1..4 row ranges");
    MRA_REQUIRE(n > 0 && n % 8 == 0 && n <= 32 * MAX_VEC * 8, "layernorm width %d unsupported (multiple of 8, <= 2048)", n);
    LnGroup grp;
    int total = 0;
    for (int i = 0; i < 4; ++i) {
        grp.seg[i] = segs[i < nseg ? i : 0];
        grp.start[i] = total;
        if (i < nseg) total += segs[i].rows;
    }
    grp.start[4] = total;
    for (int i = nseg; i <= 4; ++i) grp.start[i] = total;
    grp.n = nseg;
    const int blocks = (total + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    MRA_CHECK_CUDA(launch_pdl(layernorm_grouped_kernel, dim3(blocks), dim3(WARPS_PER_BLOCK * 32), 0, s, 1, grp, n, eps));
    return 0;
}

int launch_modality_layernorm(const void* x, int in_dtype, const float* g, const float* b, void* out, int bs, int frames,
                              int Nk, int W, int frame_major, float eps, cudaStream_t s) {
    MRA_REQUIRE(bs > 0 && frames > 0 && Nk > 0 && W > 0 && W % 8 == 0 && W <= 32 * MAX_VEC * 8,
                "modality layernorm width %d unsupported (multiple of 8, <= 2048)", W);
    MRA_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "modality layernorm: unknown input dtype %d", in_dtype);
    const int64_t tokens = static_cast<int64_t>(bs) * frames * Nk;
    const unsigned blocks = static_cast<unsigned>((tokens + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (in_dtype == 0) modality_ln_kernel<0><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, g, b, o, bs, frames, Nk, W, frame_major, eps);
    if (in_dtype == 1) modality_ln_kernel<1><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, g, b, o, bs, frames, Nk, W, frame_major, eps);
    if (in_dtype == 2) modality_ln_kernel<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, g, b, o, bs, frames, Nk, W, frame_major, eps);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_add_frame_pos(const void* x, int in_dtype, const float* pos, void* out, int bs, int frames, int n, int W,
                         cudaStream_t s) {
    MRA_REQUIRE(bs > 0 && frames > 0 && n > 0 && W > 0 && W % 8 == 0, "frame position embedding: bad shape");
    MRA_REQUIRE(in_dtype >= 0 && in_dtype <= 2, "frame position embedding: unknown input dtype %d", in_dtype);
    const int64_t tokens = static_cast<int64_t>(bs) * frames * n;
    const unsigned blocks = static_cast<unsigned>((tokens + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (in_dtype == 0) add_frame_pos_kernel<0><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, pos, o, tokens, frames, n, W);
    if (in_dtype == 1) add_frame_pos_kernel<1><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, pos, o, tokens, frames, n, W);
    if (in_dtype == 2) add_frame_pos_kernel<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, pos, o, tokens, frames, n, W);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_embed_layernorm(const float* query_embeds, int q_rows, const int32_t* ids, const void* word_emb,
                           const void* pos_emb, const float* g, const float* b, float* y32, void* y16, void* ylo, float* pre_out,
                           int rows, int Nq, int T, int H, int vocab, float eps, cudaStream_t s, const DropoutParams* drop) {
    MRA_REQUIRE(H % 8 == 0 && H <= 32 * MAX_VEC * 8, "embedding width %d unsupported", H);
    MRA_REQUIRE(T == 0 || (ids && word_emb && pos_emb), "text tokens given but ids / embedding tables are NULL");
    const int64_t total = static_cast<int64_t>(rows) * (Nq + T);
    const unsigned blocks = static_cast<unsigned>((total + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    MRA_CHECK_CUDA(launch_pdl(embed_ln_kernel, dim3(blocks), dim3(WARPS_PER_BLOCK * 32), 0, s, 1, query_embeds, q_rows, ids,
                              reinterpret_cast<const __nv_bfloat16*>(word_emb), reinterpret_cast<const __nv_bfloat16*>(pos_emb), g,
                              b, y32, reinterpret_cast<__nv_bfloat16*>(y16), reinterpret_cast<__nv_bfloat16*>(ylo), pre_out, rows,
                              Nq, T, H, vocab, eps, drop ? *drop : DropoutParams()));
    return 0;
}

int launch_build_enc_mask(const int32_t* enc_mask, float* out, int rows, int Nk, cudaStream_t s) {
    const int64_t n = static_cast<int64_t>(rows) * Nk;
    enc_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(enc_mask, out, n);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// LLM prompt assembly: one block per (segment row, frame, video) copies D bf16 (16-byte accesses) into inputs_embeds.
struct PromptSegs {
    mra_prompt_segment seg[MRA_MAX_PROMPT_SEGMENTS];
    int row_start[MRA_MAX_PROMPT_SEGMENTS + 1];   // prefix sums of rows * frames
    int n;
};
namespace {
__global__ void __launch_bounds__(256) prompt_assemble_kernel(__nv_bfloat16* __restrict__ out, int L, int D, const PromptSegs ps) {
    const int item = blockIdx.x, b = blockIdx.y;
    int s = 0;
    for (int i = 1; i < ps.n; ++i)
        if (item >= ps.row_start[i]) s = i;
    const mra_prompt_segment& sg = ps.seg[s];
    const int t = item - ps.row_start[s];
    const int f = t / sg.rows, r = t - f * sg.rows;
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(sg.src) + b * sg.src_video_stride + f * sg.src_frame_stride +
                               static_cast<int64_t>(r) * D;
    __nv_bfloat16* dst = out + (static_cast<int64_t>(b) * L + sg.dst_row + static_cast<int64_t>(f) * sg.dst_frame_rows + r) * D;
    for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8)
        *reinterpret_cast<uint4*>(dst + c) = __ldg(reinterpret_cast<const uint4*>(src + c));
}
}  // namespace

int launch_prompt_assemble(void* out, int bs, int L, int D, const mra_prompt_segment* segs, int n, cudaStream_t s) {
    MRA_REQUIRE(out && bs > 0 && L > 0 && D > 0 && D % 8 == 0, "prompt assembly: bad shape bs=%d L=%d D=%d", bs, L, D);
    MRA_REQUIRE(n >= 0 && n <= MRA_MAX_PROMPT_SEGMENTS, "prompt assembly takes at most %d segments, got %d", MRA_MAX_PROMPT_SEGMENTS, n);
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "inputs_embeds must be 16-byte aligned");
    PromptSegs ps;
    ps.n = n;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        const mra_prompt_segment& g = segs[i];
        MRA_REQUIRE(g.src && g.rows > 0 && g.frames > 0, "prompt segment %d: empty", i);
        MRA_REQUIRE((reinterpret_cast<uintptr_t>(g.src) & 15) == 0 && g.src_video_stride % 8 == 0 && g.src_frame_stride % 8 == 0,
                    "prompt segment %d: source must be 16-byte aligned", i);
        MRA_REQUIRE(g.dst_row >= 0 && g.dst_row + static_cast<int64_t>(g.frames - 1) * g.dst_frame_rows + g.rows <= L,
                    "prompt segment %d exceeds the sequence (L=%d)", i, L);
        ps.seg[i] = g;
        ps.row_start[i] = total;
        total += g.rows * g.frames;
    }
    for (int i = n; i <= MRA_MAX_PROMPT_SEGMENTS; ++i) ps.row_start[i] = total;
    if (total == 0) return 0;
    prompt_assemble_kernel<<<dim3(total, bs), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(out), L, D, ps);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_gather_last_hidden(const float* split, float* out, int rows, int Nq, int T, int H, cudaStream_t s) {
    MRA_REQUIRE(H % 4 == 0, "hidden size must be a multiple of 4");
    const int64_t n = static_cast<int64_t>(rows) * (Nq + T) * (H / 4);
    gather_last_hidden_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const float4*>(split), reinterpret_cast<float4*>(out), rows, Nq, T, H / 4);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_gather_last_hidden_split(const void* hi, const void* lo, float* out, int rows, int Nq, int T, int H, cudaStream_t s) {
    MRA_REQUIRE(H % 8 == 0, "hidden size must be a multiple of 8");
    const int64_t n = static_cast<int64_t>(rows) * (Nq + T) * (H / 8);
    gather_last_hidden_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const uint4*>(hi), reinterpret_cast<const uint4*>(lo), reinterpret_cast<float4*>(out), rows, Nq, T, H / 8);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mra
