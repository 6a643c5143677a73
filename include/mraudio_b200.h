/*
 * mraudio_b200 -- C-ABI of the B200-native (sm_100a) Q-Former / llm_proj / moment-scoring hot path of globc/mrAudio.
 *
 * The reference is pure Python and has no FFI of its own; each entry point below names the reference interface
 * (file:line under /root/reference) whose arithmetic it replaces.  A maintainer binds this library with ctypes
 * (see INTEGRATION.md); mraudio_b200/_lib.py is that binding.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; mra_last_error() returns a thread-local message.
 *     Nothing throws or exits across the boundary.
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch allocates inputs, outputs and workspace);
 *     the library owns only its handle (TMA descriptors, launch configuration).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises, the legacy default
 *     stream is never used implicitly.  A handle is not thread-safe: one per rank/stream.
 *   - activations / weights are bf16 (row-major, K contiguous); biases, LayerNorm parameters and accumulators are fp32;
 *     the residual stream is fp32 (training) or a bf16 hi + lo pair (inference, see mra_gemm_ln_split_bf16); ids and
 *     masks int32.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute entry returns an error.
 */
#ifndef MRAUDIO_B200_H
#define MRAUDIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRA_MAX_LAYERS 16
#define MRA_NUM_IOU_THDS 10

/* ---- library ------------------------------------------------------------------------------------------------ */
const char* mra_last_error(void);
int mra_version(void);
/* 0 if the current CUDA device can run the kernels (compute capability 10.x), else an error. */
int mra_device_check(void);

/* ---- Q-Former handle ---------------------------------------------------------------------------------------- */
/* Mirrors BertConfig as built by XInstructBLIP.init_Qformer (models/xinstructblip.py:615-623). */
typedef struct mra_qformer_config {
    int32_t hidden;      /* 768  */
    int32_t layers;      /* 12   */
    int32_t heads;       /* 12   (head_dim must be 64) */
    int32_t inter;       /* 3072 */
    int32_t enc_width;   /* encoder_width: 1408 video / 768 audio */
    int32_t cross_freq;  /* cross_attention_freq: 2 */
    int32_t num_query;   /* query_length: 32 */
    int32_t llm_dim;     /* llm_proj out features: 4096 (0 = no projection) */
    int32_t vocab;       /* 30523 */
    int32_t max_pos;     /* 512 */
    float ln_eps;        /* 1e-12 */
} mra_qformer_config;

/* One BertLayer.  Matrices bf16 [out, in]; vectors fp32.  Cross-attention members are NULL on layers without it. */
typedef struct mra_qformer_layer_weights {
    const void* w_qkv;  const float* b_qkv;          /* attention.self.{query,key,value} stacked: [3H, H], [3H] */
    const void* w_ao;   const float* b_ao;           /* attention.output.dense */
    const float* ln_a_g; const float* ln_a_b;        /* attention.output.LayerNorm */
    const void* w_cq;   const float* b_cq;           /* crossattention.self.query */
    const void* w_co;   const float* b_co;           /* crossattention.output.dense */
    const float* ln_c_g; const float* ln_c_b;        /* crossattention.output.LayerNorm */
    const void* w_fq1;  const float* b_fq1;          /* intermediate_query.dense [I, H] */
    const void* w_fq2;  const float* b_fq2;          /* output_query.dense [H, I] */
    const float* ln_fq_g; const float* ln_fq_b;      /* output_query.LayerNorm */
    const void* w_ft1;  const float* b_ft1;          /* intermediate.dense */
    const void* w_ft2;  const float* b_ft2;          /* output.dense */
    const float* ln_ft_g; const float* ln_ft_b;      /* output.LayerNorm */
} mra_qformer_layer_weights;

typedef struct mra_qformer_weights {
    const void* word_emb;   /* bf16 [vocab, H]    bert.embeddings.word_embeddings      (NULL for query-only) */
    const void* pos_emb;    /* bf16 [max_pos, H]  bert.embeddings.position_embeddings  (NULL for query-only) */
    const float* ln_e_g; const float* ln_e_b;        /* bert.embeddings.LayerNorm */
    /* crossattention.self.{key,value} of ALL cross layers stacked so the encoder tokens are read once:
       rows [c*2H, c*2H+H) = key of the c-th cross layer, [c*2H+H, (c+1)*2H) = its value.  bf16 [ncross*2H, W]. */
    const void* w_ckv;  const float* b_ckv;
    const void* w_proj; const float* b_proj;         /* {modality}_llm_proj: bf16 [D, H], fp32 [D] */
    mra_qformer_layer_weights layer[MRA_MAX_LAYERS];
} mra_qformer_weights;

typedef struct mra_qformer mra_qformer_t;

int mra_qformer_create(const mra_qformer_config* cfg, mra_qformer_t** out);
int mra_qformer_set_weights(mra_qformer_t* h, const mra_qformer_weights* w);
void mra_qformer_destroy(mra_qformer_t* h);

/* flags for forward */
#define MRA_FWD_SKIP_DEAD_TEXT_FFN 1u /* last layer's text FFN is never read by llm_proj (xinstructblip.py:303) */
#define MRA_FWD_SAVE_FOR_BACKWARD  2u /* keep per-layer activations in the workspace for mra_qformer_backward */

typedef struct mra_qformer_io {
    /* inputs */
    const void* enc;            /* bf16 [rows, Nk, W]   encoder_hidden_states (already through {modality}_ln) */
    const int32_t* input_ids;   /* [rows, T] or NULL when T == 0 */
    const int32_t* attn_mask;   /* [rows, Nq+T] attention_mask over queries||text: 1 = attend, 0 = padding; NULL = ones */
    const int32_t* enc_mask;    /* [rows, Nk] or NULL = all ones (the reference always passes ones, :266,275) */
    const float* query_embeds;  /* fp32 [q_rows, Nq, H], q_rows == 1 (broadcast) or rows */
    int32_t q_rows;
    int32_t rows, T, Nk;
    uint32_t flags;
    /* outputs (either may be NULL) */
    float* last_hidden;         /* fp32 [rows, Nq+T, H]  == Qformer.bert(...).last_hidden_state */
    void* llm_out;              /* bf16 [rows*Nq, D]     == llm_proj(last_hidden_state[:, :Nq]) viewed [bs, F*Nq, D] */
    /* Optional scatter of llm_out into the interleaved LLM prompt (models/xinstructblip.py:359-366: per frame
     * cue || 32 video tokens || cue || 32 audio tokens || timestamp): with llm_frames = F > 0, row (b*F + f)*Nq + q is
     * written to llm_out + b*llm_video_stride + f*llm_frame_stride + q*llm_ld (bf16 elements; llm_ld 0 = llm_dim), i.e.
     * llm_out points at the slot of (video 0, frame 0) inside inputs_embeds [bs, L, D].  Needs Nq == 32.  All 0 = dense. */
    int32_t llm_frames;
    int32_t reserved0;
    int64_t llm_ld, llm_frame_stride, llm_video_stride;
    /* Training-mode dropout (model.train() of utils/trainer.py:110; BertConfig hidden_dropout_prob = attention_probs_dropout_prob
     * = 0.1): with dropout_p > 0 -- only together with MRA_FWD_SAVE_FOR_BACKWARD -- the embeddings' LayerNorm output, the
     * attention probabilities (self and cross) and the output of every attention-output / FFN-output Linear (before the
     * residual add) are dropped with a counter-based Philox mask derived from dropout_seed (never stored: mra_qformer_backward
     * regenerates it from the same io).  Effective probability round(p * 256) / 256, see csrc/dropout.cuh.  The text length
     * T must be a multiple of 32 (pad the prompt and mask the padding) so that the TMA attention kernels cover the shape. */
    float dropout_p;
    uint32_t reserved1;
    uint64_t dropout_seed;
    /* Optional cudaEvent_t: `enc` is still being produced (e.g. copied from the host on another stream) when the call is
     * made; the forward makes its stream wait for the event in front of the first kernel that reads `enc` (the stacked
     * cross-attention K/V projection), so the embeddings and -- in a lockstep call -- the other Q-Former's projection need
     * not wait for it.  NULL = `enc` is ready in stream order. */
    void* enc_ready;
} mra_qformer_io;

/* Replaces `{modality}_Qformer.bert(input_ids, attention_mask=, query_embeds=, encoder_hidden_states=,
 * encoder_attention_mask=, return_dict=True)` (models/xinstructblip.py:286-293, 461-468) and the following
 * `{modality}_llm_proj(last_hidden_state[:, :32, :])` (:303, :475). */
size_t mra_qformer_workspace_bytes(const mra_qformer_t* h, int32_t rows, int32_t T, int32_t Nk, uint32_t flags);
int mra_qformer_forward(mra_qformer_t* h, const mra_qformer_io* io, void* workspace, size_t workspace_bytes,
                        void* stream);
/* Same for up to two Q-Formers of one batch (video + audio) in lockstep: each Linear of a layer is ONE grouped GEMM launch
 * over both (FFN_query / FFN_text are further problems of the same launch).  The handles must share the layer geometry
 * (hidden, heads, layers, intermediate, queries, cross frequency, llm_dim); encoder width, rows, T, Nk may differ.
 * The launch count / profile of the call is reported on hs[0]. */
int mra_qformer_forward_multi(int32_t n, mra_qformer_t* const* hs, const mra_qformer_io* const* ios, void* const* workspaces,
                              const size_t* workspace_bytes, void* stream);
/* number of kernels the last forward call enqueued (for bench.py's gpu_launches) */
int mra_qformer_last_launch_count(const mra_qformer_t* h);

/* ---- fine-tuning: backward of the Q-Former / projection parameters + optimizer ---------------------------------
 * Replaces the autograd + optimizer the reference runs in utils/trainer.py:129-140 for the trainable parameters (here
 * the Q-Former, query tokens and llm_proj; encoders and LLM frozen: no gradient flows into `enc`).
 *   1. mra_qformer_forward(io with MRA_FWD_SAVE_FOR_BACKWARD) keeps per-layer activations in `workspace`;
 *   2. mra_qformer_backward(same io, same workspace, d_llm = dL/d(llm_out) bf16 [rows*Nq, D]) ACCUMULATES fp32
 *      gradients into `g` (same packed layout as mra_qformer_weights: stacked q,k,v / stacked cross k,v);
 *      `reserved` must be NULL (the dgrad GEMMs read the forward's weights in place through MN-major descriptors, see
 *      mra_dgrad_bf16: no transposed weight copies exist);
 *   3. mra_adam_step: torch.optim.Adam semantics on flat fp32 buffers (utils/trainer.py:65), grads pre-multiplied by
 *      grad_scale (1 / (accum_grad_iters * world_size) after a sum all-reduce);  mra_cast_bf16 refreshes bf16 copies. */
typedef struct mra_qformer_layer_grads {
    float* w_qkv;  float* b_qkv;  float* w_ao;   float* b_ao;   float* ln_a_g;  float* ln_a_b;
    float* w_cq;   float* b_cq;   float* w_co;   float* b_co;   float* ln_c_g;  float* ln_c_b;
    float* w_fq1;  float* b_fq1;  float* w_fq2;  float* b_fq2;  float* ln_fq_g; float* ln_fq_b;
    float* w_ft1;  float* b_ft1;  float* w_ft2;  float* b_ft2;  float* ln_ft_g; float* ln_ft_b;
} mra_qformer_layer_grads;
typedef struct mra_qformer_grads {
    float* word_emb; float* pos_emb;   /* may be NULL (frozen embeddings) */
    float* ln_e_g; float* ln_e_b;
    float* w_ckv; float* b_ckv;
    float* w_proj; float* b_proj;
    float* query_tokens;               /* fp32 [q_rows, Nq, H]; may be NULL */
    mra_qformer_layer_grads layer[MRA_MAX_LAYERS];
} mra_qformer_grads;

size_t mra_qformer_backward_workspace_bytes(const mra_qformer_t* h, int32_t rows, int32_t T, int32_t Nk);
int mra_qformer_backward(mra_qformer_t* h, const mra_qformer_io* io, const void* d_llm, const void* reserved,
                         const mra_qformer_grads* g, void* workspace, size_t workspace_bytes, void* bwd_workspace,
                         size_t bwd_bytes, void* stream);
/* Overlap of the data-parallel gradient all-reduce (DistributedDataParallel, utils/trainer.py:69) with the backward:
 * events[l] (cudaEvent_t, owned by the caller; NULL entries allowed) is recorded on the backward's stream as soon as the
 * gradients of layer l -- and therefore of all layers above it -- are final, so the caller can start the all-reduce of
 * that bucket on another stream while the lower layers are still running; events[layers] (optional, n = layers + 1) is
 * recorded when the projection's gradients are final, i.e. after the first launches of the backward.  At each event the
 * WEIGHTS of those layers (of the projection) are also no longer read by this backward: the caller may run the optimizer
 * update of the bucket -- rewriting the bf16 operands in place -- behind the event.  (The stacked
 * cross-attention K/V gradients -- ONE weight-gradient GEMM over all cross layers at the end --, the embeddings and the
 * query tokens are final only when mra_qformer_backward has finished.)  n = 0 clears. */
int mra_qformer_backward_layer_events(mra_qformer_t* h, void* const* events, int32_t n);
int mra_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t step, float grad_scale, void* stream);
/* The same Adam update fused with what always follows it in the fine-tuning loop (utils/trainer.py:137-140): the bf16
 * operand copy of the updated parameters (params_bf16, may be NULL) and optimizer.zero_grad() (zero_grads != 0).
 * n must be a multiple of 4, buffers 16-byte aligned. */
int mra_adam_step_fused(float* params, float* grads, const void* reduced_grads_bf16, float* exp_avg, float* exp_avg_sq,
                        void* params_bf16, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int32_t step, float grad_scale, int32_t zero_grads, void* stream);
/* mra_adam_step_fused with the per-step scalars read from DEVICE memory -- hyper_dev[4] = {lr, 1 - beta1^step,
 * sqrt(1 - beta2^step), grad_scale}, 16-byte aligned -- so that the launch can be captured in a CUDA graph and replayed
 * while the learning-rate schedule (utils/trainer.py:127) and the step count advance.  mra_adam_hyper fills the four values
 * on the host exactly as mra_adam_step_fused derives them (bit-identical updates). */
int mra_adam_step_fused_dyn(float* params, float* grads, const void* reduced_grads_bf16, float* exp_avg, float* exp_avg_sq,
                            void* params_bf16, int64_t n, float beta1, float beta2, float eps, float weight_decay,
                            int32_t zero_grads, const float* hyper_dev, void* stream);
void mra_adam_hyper(float lr, float beta1, float beta2, int32_t step, float grad_scale, float* out4);
int mra_cast_bf16(const float* in, void* out, int64_t n, void* stream);

/* Device-side timing of the launches of mra_qformer_forward with CUDA events recorded on the caller's stream.
 *   MRA_PROFILE_DOMINANT brackets only the tensor-core GEMM launches (the dominant kernel: bench.py's roofline),
 *   MRA_PROFILE_ALL every launch.  mra_qformer_profile_read synchronises on the recorded events, returns the summed
 *   milliseconds and launch counts per category since the previous read, and clears them. */
#define MRA_PROFILE_OFF 0
#define MRA_PROFILE_DOMINANT 1
#define MRA_PROFILE_ALL 2
#define MRA_CAT_GEMM_CROSS_KV 0 /* cross-attention K/V projection of all cross layers (one GEMM) */
#define MRA_CAT_GEMM 1          /* every other Linear */
#define MRA_CAT_ATTENTION 2
#define MRA_CAT_LAYERNORM 3
#define MRA_CAT_OTHER 4
#define MRA_NUM_CATS 5
int mra_qformer_profile_mode(mra_qformer_t* h, int32_t mode);
int mra_qformer_profile_read(mra_qformer_t* h, double* ms_by_cat, int64_t* launches_by_cat);

/* ---- LLM prompt assembly (models/xinstructblip.py:342-385, 544-594) -----------------------------------------------
 * inputs_embeds [bs, L, D] bf16 is the concatenation, per video, of F frame blocks (optional enumeration tokens, cue,
 * 32 video tokens, cue, 32 audio tokens, timestamp tokens) followed by the duration and the prompt embeddings.  The
 * Q-Former tokens are written in place by llm_proj (mra_qformer_io::llm_frames); this call copies every OTHER piece
 * (embeddings of LLM tokens, produced by the caller's frozen LLM embedding table) to its slot in one launch.
 * Segment s: rows [0, rows) of src (bf16, row length D) go to dst rows [dst_row + f*dst_frame_rows, ...) of every video b,
 * for f in [0, frames); the source row block is src + b*src_video_stride + f*src_frame_stride (elements; 0 = the same
 * piece for every video / frame, as for the cues). */
typedef struct mra_prompt_segment {
    const void* src;
    int64_t src_video_stride, src_frame_stride;
    int32_t rows, frames;
    int32_t dst_row, dst_frame_rows;
} mra_prompt_segment;
#define MRA_MAX_PROMPT_SEGMENTS 16
int mra_prompt_assemble(void* inputs_embeds, int32_t bs, int32_t L, int32_t D, const mra_prompt_segment* segs, int32_t n_segs,
                        void* stream);

/* ---- building-block ops (each is also what the forward above launches; exposed for parity tests and backward) */

/* C[M,N] = epilogue(A[M,K] . W[N,K]^T): tcgen05 tensor-core GEMM, TMA-fed, fp32 accumulation in TMEM.
 *   A, W bf16 with row strides lda/ldw (elements, multiples of 8); bias fp32 [N] or NULL; residual fp32 [M,N] with
 *   stride ldr or NULL; gelu != 0 applies erf-GELU after the bias; out_fp32 selects fp32 vs bf16 C (stride ldc).
 *   Replaces torch.nn.Linear inside the Q-Former (LAVIS Qformer.py; HF port modeling_instructblip.py:499-509,549-553,
 *   586-610) and `llm_proj` (models/xinstructblip.py:707-708). */
#define MRA_GEMM_IMPL_TCGEN05 0
#define MRA_GEMM_IMPL_SIMT_DEBUG 1 /* slow CUDA-core kernel, tests only: isolates tensor-core descriptor bugs */
int mra_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                  int64_t ldr, void* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t gelu, int32_t out_fp32,
                  int32_t impl, void* stream);

/* Weight gradient of a Linear:  dW[n_out, k_in] (+)= dY[n, n_out]^T . X[n, k_in]  on the same tcgen05 kernel, reading dY and X
 * as they lie in memory (row-major, the reduction index n is the row index) through MN-major shared-memory descriptors:
 * no transposed copies.  dY, X bf16; dW fp32 (row stride ldw); accumulate != 0 adds to dW. */
int mra_wgrad_bf16(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t ldw, int32_t n, int32_t n_out,
                   int32_t k_in, int32_t accumulate, void* stream);

/* Data gradient of a Linear:  dX[n, k_in] = dY[n, n_out] . W[n_out, k_in] (+ residual)  on the same kernel, reading the
 * forward's weight matrix W as it lies (MN-major descriptor for the second operand): no transposed weight copies.
 * dY, W bf16; dX bf16 or fp32 (out_fp32); residual fp32 [n, k_in] or NULL (only with fp32 output; may alias dX). */
int mra_dgrad_bf16(const void* dY, int64_t ldy, const void* W, int64_t ldw, const float* residual, int64_t ldr, void* dX,
                   int64_t ldx, int32_t n, int32_t n_out, int32_t k_in, int32_t out_fp32, void* stream);

/* Fused  y = LayerNorm(A . W^T + bias + residual) * gamma + beta  for N == 768: Linear + residual add + post-LayerNorm of
 * BertSelfOutput / BertOutput (HF port modeling_instructblip.py:549-553, 606-610) in one kernel; the pre-LayerNorm sums
 * never leave TMEM / registers.  A bf16 [M, K], W bf16 [768, K], residual fp32 [M, 768]; writes y32 (fp32) and y16 (bf16).
 * A 6-CTA thread-block cluster (3 column slices of 256 x a CTA pair running tcgen05.mma.cta_group::2, M = 256) owns each
 * 256-row block and combines the per-row statistics through distributed shared memory (csrc/gemm_ln.cu). */
int mra_gemm_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                     int64_t ldr, const float* gamma, const float* beta, float* y32, int64_t ldy32, void* y16, int64_t ldy16,
                     int32_t M, int32_t N, int32_t K, float eps, void* stream);

/* Same with the residual stream as a PAIR of bf16 tensors (value = hi + lo, hi = bf16(value), lo = bf16(value - hi): 16
 * mantissa bits): res_hi / res_lo in (row stride ldr), y_hi / y_lo out (row stride ldy); y_hi is what the next Linear reads
 * as its A operand.  4 + 4 instead of 4 + 6 bytes per post-LayerNorm element; this is the form mra_qformer_forward uses. */
int mra_gemm_ln_split_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* res_hi,
                           const void* res_lo, int64_t ldr, const float* gamma, const float* beta, void* y_hi, void* y_lo,
                           int64_t ldy, int32_t M, int32_t N, int32_t K, float eps, void* stream);

/* Tuning aid: force the output-tile width of the tcgen05 GEMM (128, 192 or 256; 0 = automatic choice). */
int mra_gemm_tile_override(int32_t bn);

/* Tuning / test aid, applied when every problem has >= 4 row blocks: 3 (default) pairs CTAs into 2-CTA clusters that run
 * one tcgen05.mma.cta_group::2 per K step (M = 256, half of the W slab per CTA); 2 pairs them sharing the W slab by TMA
 * multicast; 1 never pairs. */
int mra_gemm_cluster_override(int32_t cm);

/* Fused multi-head attention core, head_dim 64: O = softmax(Q K^T / 8 + mask) V, per (row, head).
 *   q: bf16, row `qrow(r, i)` at q + qrow*ldq + head*64;  k, v likewise with ldk/ldv; o bf16 with ldo.
 *   Row addressing: "split" layout used by the forward: query tokens of all rows first, then text tokens:
 *     index(r, i) = i < nq_split ? r*nq_split + i : rows*nq_split + r*(S - nq_split) + (i - nq_split)
 *   Self-attention: Sq = Sk = S, q/k/v share the index function.  Cross-attention: nq_split = Sq (dense), keys dense
 *   (index r*Sk + j).  add_mask fp32 [rows, Sk] added to the scaled scores, or NULL.
 *   Replaces BertSelfAttention core (HF port modeling_instructblip.py:512-536). */
int mra_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                  int64_t ldo, const float* add_mask, int32_t rows, int32_t heads, int32_t Sq, int32_t Sk,
                  int32_t nq_split, int32_t kv_dense, void* stream);

/* The same core on HEAD-MAJOR operands, the layout the inference forward uses between the QKV / cross-K/V Linear and
 * the attention core: head h of q starts at q + h*hsq and its token rows are ldq elements apart (likewise k, v), so with
 * ldq = 64 and hsq = tokens*64 every (row, head) tile is one contiguous run of bytes.  hs* = 0 means 64 (heads side by
 * side in a row: exactly mra_attention).  All strides in elements, multiples of 8. */
int mra_attention_strided(const void* q, int64_t ldq, int64_t hsq, const void* k, int64_t ldk, int64_t hsk, const void* v,
                          int64_t ldv, int64_t hsv, void* o, int64_t ldo, const float* add_mask, int32_t rows,
                          int32_t heads, int32_t Sq, int32_t Sk, int32_t nq_split, int32_t kv_dense, void* stream);

/* C = A W^T + bias written head-major: bf16 C[N/64][M][64], i.e. column n of output row m at C + ((n/64)*M + m)*64 + n%64
 * (N % 64 == 0).  Producer side of mra_attention_strided: the fused QKV projection (HF port :499-510) and the stacked
 * cross-attention K/V projection write this form in the inference forward. */
int mra_gemm_head_major_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int32_t M,
                             int32_t N, int32_t K, void* stream);

/* ctx = SelfAttention(x W_qkv^T + b_qkv): the fused Q/K/V Linear of a Q-Former block and its attention core in ONE kernel
 * (Q, K, V never reach HBM), for rows of 32 query + 32 text tokens in the split layout and heads of 64:
 *   x bf16 [rows*64, K] (the 32 query tokens of every row first, then the 32 text tokens of every row), w bf16
 *   [3*heads*64, K] (Q rows, K rows, V rows), bias fp32 [3*heads*64] or NULL, add_mask fp32 [rows, 64] or NULL,
 *   ctx bf16 [rows*64, heads*64] (row order of x, 16-byte aligned rows).  Bit-equal to mra_gemm_bf16 followed by
 *   mra_attention.  Replaces BertSelfAttention query/key/value + core (HF port modeling_instructblip.py:499-538). */
int mra_qkv_attention_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* add_mask,
                           void* ctx, int64_t ldo, int32_t rows, int32_t heads, int32_t K, void* stream);

/* Test aid: generic != 0 forces the register-staged generic attention kernel instead of the TMA-pipelined one (which
 * covers Sq <= 256, Sk <= 4096 and split points that are multiples of 32; other shapes always use the generic one). */
int mra_attention_impl_override(int32_t generic);

/* y = LayerNorm(x) * gamma + beta over the last dim (n), fp32 statistics; writes fp32 and/or bf16 copies.
 *   Replaces the LayerNorm in BertSelfOutput/BertOutput (HF port :549-553, :606-610), eps 1e-12. */
int mra_layernorm(const float* x, const float* gamma, const float* beta, float* y32, void* y16, int32_t rows, int32_t n,
                  float eps, void* stream);

/* {modality}_ln: fp32-upcast LayerNorm of encoder tokens (models/xinstructblip.py:822-828, applied :265,274) fused
 * with the frame fold + batch-major reorder of :280-285.  x: [F, bs, Nk, W] (frame_major != 0: the reference's list
 * of per-frame tensors) or [bs*F, Nk, W]; in_dtype 0 = fp32, 1 = bf16, 2 = fp16.  out: bf16 [bs*F, Nk, W]. */
int mra_modality_layernorm(const void* x, int32_t in_dtype, const float* gamma, const float* beta, void* out,
                           int32_t bs, int32_t frames, int32_t Nk, int32_t W, int32_t frame_major, float eps,
                           void* stream);

/* Video-LLaMA-v1-style frame position embedding: out[b,f,t,:] = bf16(x[b,f,t,:] + pos[f,:]), x: [bs, frames, n, W]
 * (in_dtype as above), pos fp32 [frames, W].  The result viewed [bs, frames*n, W] is the key/value input of the video
 * Q-Former.  (models/videollama.py:1-25 only wraps the `videollama2` package; this arithmetic is the public
 * Video-LLaMA v1 design named by BASELINE.json config 3 -- parity unpinned, see DESIGN.md.) */
int mra_add_frame_position(const void* x, int32_t in_dtype, const float* pos, void* out, int32_t bs, int32_t frames,
                           int32_t n, int32_t W, void* stream);

/* ---- moment-retrieval scorer ----------------------------------------------------------------------------------
 * One thread per query.  Replaces compute_average_precision_detection (eval/mr_utils.py:89-171), the per-query part
 * of compute_mr_r1 (eval/mr_eval.py:97-131) and the IoU helpers (eval/mr_utils.py:16-67), in fp64 with numpy's
 * operation order, nan semantics and argsort tie order.
 *   pred [Q, Pmax, 2] f64, n_pred [Q] (>= 1); gt [Q, Gmax, 2] f64, n_gt [Q] (>= 1); thds [10] f64.
 *   out_ap [Q, 10] f64; out_iou [Q] f64 (top-1 paired IoU); out_invalid [Q] u8: bit 0 = -1 in the top-1 window;
 *   bit 1 = "tie-ambiguous": some prediction has exactly the same IoU (>= 0.5) with two different GT windows, the one case
 *   where the reference's result depends on the (platform-dependent, unstable for >= 4 elements on AVX-512 numpy builds)
 *   tie order of `tiou_arr.argsort()[::-1]` (eval/mr_utils.py:147); the kernel uses the stable order.
 *   Limits: Pmax <= 256, Gmax <= 64. */
int mra_mr_score(const double* pred, const int32_t* n_pred, const double* gt, const int32_t* n_gt, const double* thds,
                 int32_t Q, int32_t Pmax, int32_t Gmax, double* out_ap, double* out_iou, uint8_t* out_invalid,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRAUDIO_B200_H */
