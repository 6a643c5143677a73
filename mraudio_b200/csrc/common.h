// Internal declarations shared by the translation units of libmraudio_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "../../include/mraudio_b200.h"
#include "dropout.cuh"

namespace mra {

void set_error(const char* fmt, ...);

#define MRA_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            ::mra::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

#define MRA_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            ::mra::set_error(__VA_ARGS__); \
            return 2;                   \
        }                               \
    } while (0)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining, runs
// its prologue (barrier init, TMEM allocation, descriptor prefetch), and blocks in ptx::griddep_wait() until the
// predecessor has completed and flushed.  EVERY kernel launched through this helper must call ptx::griddep_wait()
// before touching global memory.
// MRA_PDL=0 disables (A/B).  Round 1 measured no gain (only the unpaired kernels took part); with every kernel of the chain
// taking part -- the 2-CTA GEMMs and the fused GEMM+LayerNorm included -- see profiles/r02_NOTES.md.
inline bool pdl_enabled() {
    static const bool enabled = [] { const char* e = getenv("MRA_PDL"); return e == nullptr || atoi(e) != 0; }();
    return enabled;
}
// cluster_x > 1: launch with that (1-D) thread-block cluster size
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cluster_x > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster_x; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int device_check();
int sm_count();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize [, PreferredSharedMemoryCarveout = max]) once per (kernel, device):
// the attribute is per device, and the cache is guarded so that handles on different threads / devices are safe
int ensure_smem_attr(const void* kernel, int bytes, bool max_carveout = false);

// ---- launchers (each returns 0 / error and bumps *launches when non-null) ----
struct GemmArgs {
    const void* A; int64_t lda;
    const void* W; int64_t ldw;
    const float* bias;
    const float* residual; int64_t ldr;
    void* C; int64_t ldc;
    int M, N, K;
    int gelu;
    int out_fp32;
    // tn != 0: A is [K, M] and W is [K, N] (the reduction index is the row index of both, row strides lda / ldw):
    // C[M, N] = A^T W  -- the weight-gradient form dW = dY^T X; needs out_fp32, no GELU
    int tn = 0;
    // c_frames > 0: C is not one [M, N] matrix but a scatter target: output row r = (b * c_frames + f) * 32 + q goes to
    // C + b * c_batch_stride + f * c_frame_stride + q * ldc (elements) -- llm_proj writing each frame's 32 query tokens
    // into their slot of the interleaved LLM prompt (models/xinstructblip.py:359-366).  bf16 output, M % 32 == 0.
    int c_frames = 0;
    int64_t c_frame_stride = 0, c_batch_stride = 0;
    // c_head_major != 0: C is [N / 64][M][64] bf16 (slot n / 64 of output row m at C + ((n / 64) * M + m) * 64): the layout
    // the attention kernel reads Q / K / V heads from as contiguous tiles.  N % 64 == 0, ldc ignored.
    int c_head_major = 0;
    // (set by the launcher) split of the K range into separate work items + TMA reduce-add epilogue, see gemm.cu
    int ksplit = 1, reduce_add = 0;
    // training-mode dropout on (A W^T + bias) BEFORE the residual is added (BertSelfOutput / BertOutput): needs the fp32
    // output + residual epilogue; drop_row0 = index of this problem's first row in the split token layout
    DropoutParams drop;
    int drop_row0 = 0;
};
int launch_gemm_tc(const GemmArgs& a, cudaStream_t s);
int launch_gemm_tc_grouped(const GemmArgs* a, int n, cudaStream_t s);   // n <= 4 problems sharing N, K, epilogue kind
int launch_gemm_simt(const GemmArgs& a, cudaStream_t s);
// cached TMA descriptor of a row-major [rows, cols] matrix (2- or 4-byte elements), box {box_cols, box_rows}, 128B swizzle
int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, int elem_bytes,
                   CUtensorMap* out);
// bf16 [batch][frames][32][cols] with element strides (batch_stride, frame_stride, ld, 1); box {64 cols, 32 rows, 1, 1}
int get_tensor_map_scatter(const void* ptr, int64_t batch, int64_t frames, int64_t cols, int64_t ld, int64_t frame_stride,
                           int64_t batch_stride, CUtensorMap* out);
int get_tensor_map_head_major(const void* ptr, int64_t rows, int64_t cols, CUtensorMap* out);
void set_gemm_tile_override(int bn);
void set_gemm_cluster_override(int cm);   // 1 = never pair CTAs, 2 = pair along M when possible (default)

// fused  y = LayerNorm(A W^T + bias + residual) * gamma + beta,  N = 768  (gemm_ln.cu)
struct GemmLnArgs {
    const void* A; int64_t lda;        // bf16 [M, K]
    const void* W; int64_t ldw;        // bf16 [768, K]
    const float* bias;                 // fp32 [768] or NULL
    const void* residual; int64_t ldr;    // fp32 [M, 768]  (bf16 hi part in split form)
    const float* gamma; const float* beta;
    float* y32; int64_t ldy32;         // fp32 [M, 768]
    void* y16; int64_t ldy16;          // bf16 [M, 768]
    int M, K;
    // split residual stream (both non-NULL selects it): `residual` then points at the bf16 hi part (row stride ldr) and
    // res_lo at the bf16 lo part; the outputs are y16 = hi and y_lo = lo (row stride ldy16), y32 is unused
    const void* res_lo = nullptr;
    void* y_lo = nullptr;
};
int launch_gemm_ln_grouped(const GemmLnArgs* a, int n, float eps, cudaStream_t s);

struct AttnArgs {
    const void* q; int64_t ldq;
    const void* k; int64_t ldk;
    const void* v; int64_t ldv;
    void* o; int64_t ldo;
    const float* add_mask;
    int rows, heads, Sq, Sk, nq_split, kv_dense;
    DropoutParams drop;   // training-mode dropout on the attention probabilities (TMA kernel only)
    // head strides in elements; 0 = the heads sit side by side in a row (head h at column 64 h).  Head-major operands
    // ([head][token][64], written by the GEMM epilogue's c_head_major form): ld = 64, head stride = tokens * 64.
    int64_t hsq = 0, hsk = 0, hsv = 0;
};
int launch_attention(const AttnArgs& a, cudaStream_t s);
int launch_attention_pair(const AttnArgs* a, int n, cudaStream_t s, int* launches);   // n <= 2, one launch when possible
void set_attention_impl_override(int generic);

// fused Q/K/V projection + self-attention core for rows of 32 query + 32 text tokens in the split layout (qkv_attn.cu)
struct QkvAttnArgs {
    const void* x; int64_t ldx;      // bf16 [rows * 64, K]: the query tokens of all rows, then the text tokens of all rows
    const void* w; int64_t ldw;      // bf16 [3 * heads * 64, K]: Q rows, K rows, V rows (the fused QKV Linear)
    const float* bias;               // fp32 [3 * heads * 64] or NULL
    void* ctx; int64_t ldo;          // bf16 [rows * 64, heads * 64] out, same row order as x
    const float* add_mask;           // fp32 [rows, 64] added to the scaled scores, or NULL
    int rows, heads, K;
};
int launch_qkv_attention(const QkvAttnArgs* a, int n, cudaStream_t s);   // n <= 2 problems in one launch

int launch_layernorm(const float* x, const float* g, const float* b, float* y32, void* y16, int rows, int n, float eps,
                     cudaStream_t s);
struct LnSegment {
    const float* x; const float* gamma; const float* beta;
    float* y32; void* y16;
    int rows;
};
int launch_layernorm_grouped(const LnSegment* segs, int nseg, int n, float eps, cudaStream_t s);
int launch_modality_layernorm(const void* x, int in_dtype, const float* g, const float* b, void* out, int bs, int frames,
                              int Nk, int W, int frame_major, float eps, cudaStream_t s);
int launch_add_frame_pos(const void* x, int in_dtype, const float* pos, void* out, int bs, int frames, int n, int W,
                         cudaStream_t s);
// embeddings: LN(cat(query_embeds, word_emb[ids] + pos_emb[:T])) written in the split layout (queries first)
int launch_embed_layernorm(const float* query_embeds, int q_rows, const int32_t* ids, const void* word_emb,
                           const void* pos_emb, const float* g, const float* b, float* y32, void* y16, void* ylo, float* pre_out,
                           int rows, int Nq, int T, int H, int vocab, float eps, cudaStream_t s,
                           const DropoutParams* drop = nullptr);
// additive masks (LAVIS get_extended_attention_mask): out[r, j] = (1 - mask[r, j]) * -10000
int launch_build_enc_mask(const int32_t* enc_mask, float* out, int rows, int Nk, cudaStream_t s);
// split layout [queries ; text] fp32 -> interleaved [rows, Nq+T, H] fp32
int launch_prompt_assemble(void* out, int bs, int L, int D, const mra_prompt_segment* segs, int n, cudaStream_t s);
int launch_gather_last_hidden(const float* split, float* out, int rows, int Nq, int T, int H, cudaStream_t s);
int launch_gather_last_hidden_split(const void* hi, const void* lo, float* out, int rows, int Nq, int T, int H, cudaStream_t s);

// ---- backward / optimizer (backward.cu)
struct AttnBwdArgs {
    const void* q; int64_t ldq;
    const void* k; int64_t ldk;
    const void* v; int64_t ldv;
    const void* d_o; int64_t ldo;
    void* dq; int64_t lddq;
    void* dk; int64_t lddk;
    void* dv; int64_t lddv;
    const float* add_mask;
    int rows, heads, Sq, Sk, nq_split, kv_dense;
    const void* o = nullptr; int64_t ldof = 0;   // forward output of the same attention (enables the tensor-core kernel)
    // optional bias gradients of the q / k / v projections (fp32 [heads * 64], accumulated): column sums of dq / dk / dv over
    // all tokens; the tensor-core kernel produces them in the same pass
    float* db_q = nullptr; float* db_k = nullptr; float* db_v = nullptr;
    int64_t n_q_tokens = 0, n_k_tokens = 0;      // total token rows of dq and of dk / dv (for the fallback column-sum pass)
    DropoutParams drop;   // the forward's attention-probability dropout (tensor-core kernel only)
};
int launch_attention_bwd(const AttnBwdArgs& a, cudaStream_t s);
int launch_transpose(const void* in, int64_t ld_in, void* out, int64_t ld_out, int R, int C, float* colsum, cudaStream_t s);
int launch_colsum(const void* in, int64_t ld, int R, int C, float* colsum, cudaStream_t s);
// drop_out: dropout sat between the producing Linear and the residual add (dx16 and dbias get the masked gradient, dx32 --
// the residual path -- does not); drop_in: dropout sat on the LayerNorm's OUTPUT (embeddings): dy is masked on load.
// row0 = index of the first row in the split token layout.
int launch_ln_bwd(const float* dy, const float* pre, const float* gamma, float* dx32, void* dx16, float* dgamma, float* dbeta,
                  float* dbias, int rows, int n, float eps, cudaStream_t s, const DropoutParams* drop_out = nullptr,
                  const DropoutParams* drop_in = nullptr, int row0 = 0);
int launch_gelu_bwd_colsum(const void* z, const void* dy, void* dz, float* colsum, int rows, int cols, cudaStream_t s);
int launch_gelu_fwd(const void* z, void* out, int64_t n, cudaStream_t s);
int launch_gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, cudaStream_t s);
int launch_embed_bwd(const float* d_emb, const int32_t* ids, float* d_query, int q_rows, float* d_word, float* d_pos, int rows,
                     int Nq, int T, int H, int vocab, cudaStream_t s);
int launch_adam_fused(float* p, float* g, const void* g16, float* m, float* v, void* p16, int64_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_scale, int zero_grad, cudaStream_t s, const float* dyn = nullptr);
// {lr, 1 - beta1^step, sqrt(1 - beta2^step), grad_scale}: what launch_adam_fused derives on the host, for the `dyn` form
void adam_hyper(float lr, float beta1, float beta2, int step, float grad_scale, float* out4);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, float grad_scale, cudaStream_t s);
int launch_cast_bf16(const float* in, void* out, int64_t n, cudaStream_t s);

int launch_mr_score(const double* pred, const int32_t* n_pred, const double* gt, const int32_t* n_gt, const double* thds,
                    int Q, int Pmax, int Gmax, double* out_ap, double* out_iou, uint8_t* out_invalid, cudaStream_t s);

}  // namespace mra
