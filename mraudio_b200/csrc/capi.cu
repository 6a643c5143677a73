// extern "C" surface of libmraudio_b200.so (see include/mraudio_b200.h) + process-wide helpers.
#include <stdarg.h>

#include <mutex>
#include <set>
#include <utility>

#include "common.h"

namespace mra {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_check() {
    static int cached = -1;  // 0 ok, >0 error code
    if (cached == 0) return 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device available (%s): mraudio_b200 has no CPU fallback", cudaGetErrorString(e));
        return 3;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
        return 3;
    }
    if (prop.major != 10) {
        set_error("device '%s' is sm_%d%d; mraudio_b200 kernels are built for sm_100a (B200) only", prop.name, prop.major, prop.minor);
        return 3;
    }
    cached = 0;
    return 0;
}

int ensure_smem_attr(const void* kernel, int bytes, bool max_carveout) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    MRA_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kernel, dev})) return 0;
    MRA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (max_carveout)
        MRA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    done.insert({kernel, dev});
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace mra

using namespace mra;

extern "C" const char* mra_last_error(void) { return g_err; }
extern "C" int mra_version(void) { return 100; }
extern "C" int mra_device_check(void) { return device_check(); }

extern "C" int mra_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                             int64_t ldr, void* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t gelu,
                             int32_t out_fp32, int32_t impl, void* stream) {
    MRA_REQUIRE(A && W && C, "mra_gemm_bf16: NULL operand");
    if (int e = device_check()) return e;
    GemmArgs a{A, lda, W, ldw, bias, residual, ldr, C, ldc, M, N, K, gelu, out_fp32};
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (impl == MRA_GEMM_IMPL_SIMT_DEBUG) return launch_gemm_simt(a, s);
    MRA_REQUIRE(impl == MRA_GEMM_IMPL_TCGEN05, "mra_gemm_bf16: unknown impl %d", impl);
    return launch_gemm_tc(a, s);
}

extern "C" int mra_wgrad_bf16(const void* dY, int64_t ldy, const void* X, int64_t ldx, float* dW, int64_t ldw, int32_t n, int32_t n_out,
                              int32_t k_in, int32_t accumulate, void* stream) {
    MRA_REQUIRE(dY && X && dW, "mra_wgrad_bf16: NULL operand");
    if (int e = device_check()) return e;
    GemmArgs a{dY, ldy, X, ldx, nullptr, accumulate ? dW : nullptr, ldw, dW, ldw, n_out, k_in, n, 0, 1};
    a.tn = 1;
    return launch_gemm_tc(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_dgrad_bf16(const void* dY, int64_t ldy, const void* W, int64_t ldw, const float* residual, int64_t ldr, void* dX,
                              int64_t ldx, int32_t n, int32_t n_out, int32_t k_in, int32_t out_fp32, void* stream) {
    MRA_REQUIRE(dY && W && dX, "mra_dgrad_bf16: NULL operand");
    if (int e = device_check()) return e;
    GemmArgs a{dY, ldy, W, ldw, nullptr, residual, ldr, dX, ldx, n, k_in, n_out, 0, out_fp32};
    a.tn = 2;
    return launch_gemm_tc(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_gemm_ln_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const float* residual,
                                int64_t ldr, const float* gamma, const float* beta, float* y32, int64_t ldy32, void* y16,
                                int64_t ldy16, int32_t M, int32_t N, int32_t K, float eps, void* stream) {
    MRA_REQUIRE(N == 768, "mra_gemm_ln_bf16: the fused GEMM + LayerNorm kernel is built for N = 768 (Q-Former hidden size), got %d", N);
    if (int e = device_check()) return e;
    GemmLnArgs a{A, lda, W, ldw, bias, residual, ldr, gamma, beta, y32, ldy32, y16, ldy16, M, K};
    return launch_gemm_ln_grouped(&a, 1, eps, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_gemm_ln_split_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* res_hi,
                                      const void* res_lo, int64_t ldr, const float* gamma, const float* beta, void* y_hi,
                                      void* y_lo, int64_t ldy, int32_t M, int32_t N, int32_t K, float eps, void* stream) {
    MRA_REQUIRE(N == 768, "mra_gemm_ln_split_bf16: the fused GEMM + LayerNorm kernel is built for N = 768 (Q-Former hidden size), got %d", N);
    MRA_REQUIRE(res_hi && res_lo && y_hi && y_lo, "mra_gemm_ln_split_bf16: NULL residual / output");
    if (int e = device_check()) return e;
    GemmLnArgs a{A, lda, W, ldw, bias, res_hi, ldr, gamma, beta, nullptr, ldy, y_hi, ldy, M, K};
    a.res_lo = res_lo;
    a.y_lo = y_lo;
    return launch_gemm_ln_grouped(&a, 1, eps, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_gemm_tile_override(int32_t bn) {
    MRA_REQUIRE(bn == 0 || bn == 128 || bn == 192 || bn == 256, "tile width override must be 0 (auto), 128, 192 or 256");
    set_gemm_tile_override(bn);
    return 0;
}

extern "C" int mra_gemm_cluster_override(int32_t cm) {
    MRA_REQUIRE(cm >= 1 && cm <= 3, "cluster override must be 1 (single CTA), 2 (pair, W multicast) or 3 (pair, 2-CTA MMA)");
    set_gemm_cluster_override(cm);
    return 0;
}

extern "C" int mra_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                             int64_t ldo, const float* add_mask, int32_t rows, int32_t heads, int32_t Sq, int32_t Sk,
                             int32_t nq_split, int32_t kv_dense, void* stream) {
    MRA_REQUIRE(q && k && v && o, "mra_attention: NULL operand");
    if (int e = device_check()) return e;
    AttnArgs a{q, ldq, k, ldk, v, ldv, o, ldo, add_mask, rows, heads, Sq, Sk, nq_split, kv_dense};
    return launch_attention(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_attention_strided(const void* q, int64_t ldq, int64_t hsq, const void* k, int64_t ldk, int64_t hsk, const void* v,
                                     int64_t ldv, int64_t hsv, void* o, int64_t ldo, const float* add_mask, int32_t rows,
                                     int32_t heads, int32_t Sq, int32_t Sk, int32_t nq_split, int32_t kv_dense, void* stream) {
    MRA_REQUIRE(q && k && v && o, "mra_attention_strided: NULL operand");
    if (int e = device_check()) return e;
    AttnArgs a{q, ldq, k, ldk, v, ldv, o, ldo, add_mask, rows, heads, Sq, Sk, nq_split, kv_dense};
    a.hsq = hsq; a.hsk = hsk; a.hsv = hsv;
    return launch_attention(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_gemm_head_major_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int32_t M,
                                        int32_t N, int32_t K, void* stream) {
    MRA_REQUIRE(A && W && C, "mra_gemm_head_major_bf16: NULL operand");
    if (int e = device_check()) return e;
    GemmArgs a{A, lda, W, ldw, bias, nullptr, 0, C, N, M, N, K, 0, 0};
    a.c_head_major = 1;
    return launch_gemm_tc(a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_qkv_attention_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, const float* add_mask,
                                      void* ctx, int64_t ldo, int32_t rows, int32_t heads, int32_t K, void* stream) {
    MRA_REQUIRE(x && w && ctx, "mra_qkv_attention_bf16: NULL operand");
    if (int e = device_check()) return e;
    QkvAttnArgs a{x, ldx, w, ldw, bias, ctx, ldo, add_mask, rows, heads, K};
    return launch_qkv_attention(&a, 1, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_attention_impl_override(int32_t generic) {
    set_attention_impl_override(generic);
    return 0;
}

extern "C" int mra_layernorm(const float* x, const float* gamma, const float* beta, float* y32, void* y16, int32_t rows,
                             int32_t n, float eps, void* stream) {
    MRA_REQUIRE(x && gamma && beta && (y32 || y16), "mra_layernorm: NULL operand");
    if (int e = device_check()) return e;
    return launch_layernorm(x, gamma, beta, y32, y16, rows, n, eps, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_modality_layernorm(const void* x, int32_t in_dtype, const float* gamma, const float* beta, void* out,
                                      int32_t bs, int32_t frames, int32_t Nk, int32_t W, int32_t frame_major, float eps,
                                      void* stream) {
    MRA_REQUIRE(x && gamma && beta && out, "mra_modality_layernorm: NULL operand");
    if (int e = device_check()) return e;
    return launch_modality_layernorm(x, in_dtype, gamma, beta, out, bs, frames, Nk, W, frame_major, eps,
                                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_add_frame_position(const void* x, int32_t in_dtype, const float* pos, void* out, int32_t bs, int32_t frames,
                                      int32_t n, int32_t W, void* stream) {
    MRA_REQUIRE(x && pos && out, "mra_add_frame_position: NULL operand");
    if (int e = device_check()) return e;
    return launch_add_frame_pos(x, in_dtype, pos, out, bs, frames, n, W, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_prompt_assemble(void* inputs_embeds, int32_t bs, int32_t L, int32_t D, const mra_prompt_segment* segs,
                                   int32_t n_segs, void* stream) {
    MRA_REQUIRE(inputs_embeds && (segs || n_segs == 0), "mra_prompt_assemble: NULL argument");
    if (int e = device_check()) return e;
    return launch_prompt_assemble(inputs_embeds, bs, L, D, segs, n_segs, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_mr_score(const double* pred, const int32_t* n_pred, const double* gt, const int32_t* n_gt,
                            const double* thds, int32_t Q, int32_t Pmax, int32_t Gmax, double* out_ap, double* out_iou,
                            uint8_t* out_invalid, void* stream) {
    MRA_REQUIRE(pred && n_pred && gt && n_gt && thds && out_ap && out_iou && out_invalid, "mra_mr_score: NULL operand");
    if (int e = device_check()) return e;
    return launch_mr_score(pred, n_pred, gt, n_gt, thds, Q, Pmax, Gmax, out_ap, out_iou, out_invalid,
                           reinterpret_cast<cudaStream_t>(stream));
}
