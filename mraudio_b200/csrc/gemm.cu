// C[M,N] = epilogue(A[M,K] . W[N,K]^T)  -- the Linear layers of the Q-Former and llm_proj.
//
// Blackwell-native design (sm_100a):
//   * persistent kernel, one CTA per SM, static round-robin over 128 x BN output tiles (n fastest so that the CTAs
//     running concurrently share the same A rows through L2);
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128B-swizzled 64-wide K slabs, STAGES-deep mbarrier ring),
//     warp 1 = MMA issuer (one thread, tcgen05.mma kind::f16, M=128, N=BN, K=16, fp32 accumulators in TMEM),
//     warps 2..5 = epilogue (tcgen05.ld 32x32b, bias / erf-GELU / fp32 residual fused, vectorised stores);
//   * two TMEM accumulator stages (2*BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
// Reference arithmetic replaced: torch.nn.Linear (+ GELU, + residual add) in LAVIS Qformer.py / HF port
// modeling_instructblip.py:499-509,549-553,586-610 and llm_proj (models/xinstructblip.py:707-708).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

namespace mra {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int NUM_THREADS = 192;

struct EpiParams {
    const float* bias;
    const float* residual;
    int64_t ldr;
    void* C;
    int64_t ldc;
    int M, N, K;
};

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;  // + tmem ptr + 1024B alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int STAGES, bool GELU, bool OUT_F32>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const EpiParams p) {
    using L = SmemLayout<BN, STAGES>;
    constexpr uint32_t TMEM_COLS = 2 * BN;  // 256 or 512: power of two >= 32
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.M + BM - 1) / BM;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 4);  // one arrival per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    ptx::tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(sB + stage * L::B_STAGE_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
                    const uint64_t b_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + stage * L::B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the >>4 address field
                        ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tfull_bar[acc]);  // accumulator complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..5
        const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const int row = m_blk * BM + quad * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(t_row + c, r);
                ptx::tmem_ld_wait();
                const int col0 = n_blk * BN + c;
                if (row_ok && col0 < p.N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < p.N) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        }
                    }
                    if (GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    }
                    if (p.residual != nullptr) {
                        const float* rp = p.residual + static_cast<int64_t>(row) * p.ldr + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < p.N) {
                                const float4 x = *reinterpret_cast<const float4*>(rp + j);
                                v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
                            }
                        }
                    }
                    if (OUT_F32) {
                        float* cp = reinterpret_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < p.N)
                                *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    } else {
                        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(row) * p.ldc + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            if (col0 + j < p.N) {
                                uint4 o;
                                o.x = ptx::pack_bf16x2(v[j], v[j + 1]);
                                o.y = ptx::pack_bf16x2(v[j + 2], v[j + 3]);
                                o.z = ptx::pack_bf16x2(v[j + 4], v[j + 5]);
                                o.w = ptx::pack_bf16x2(v[j + 6], v[j + 7]);
                                *reinterpret_cast<uint4*>(cp + j) = o;
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Debug / test-only CUDA-core kernel with the same contract (isolates descriptor bugs in the tensor-core path).
template <bool GELU, bool OUT_F32>
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda, const __nv_bfloat16* __restrict__ W,
                                 int64_t ldw, const EpiParams p) {
    __shared__ float sa[16][17];
    __shared__ float sw[16][17];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * 16 + ty;
    const int col = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        const int ar = blockIdx.y * 16 + ty, ak = k0 + tx;
        sa[ty][tx] = (ar < p.M && ak < p.K) ? __bfloat162float(A[ar * lda + ak]) : 0.f;
        const int wr = blockIdx.x * 16 + ty;
        sw[ty][tx] = (wr < p.N && ak < p.K) ? __bfloat162float(W[wr * ldw + ak]) : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sw[tx][k], acc);
        __syncthreads();
    }
    if (row < p.M && col < p.N) {
        if (p.bias) acc += p.bias[col];
        if (GELU) acc = gelu_erf(acc);
        if (p.residual) acc += p.residual[static_cast<int64_t>(row) * p.ldr + col];
        if (OUT_F32)
            reinterpret_cast<float*>(p.C)[static_cast<int64_t>(row) * p.ldc + col] = acc;
        else
            reinterpret_cast<__nv_bfloat16*>(p.C)[static_cast<int64_t>(row) * p.ldc + col] = __float2bfloat16_rn(acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA descriptor creation (driver entry point fetched through the runtime: no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    int64_t rows, cols, ld;
    int box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        h = h * 1000003u ^ static_cast<size_t>(k.cols);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.box_rows);
        return h;
    }
};

// bf16 row-major [rows, cols] with row stride ld (elements); box = {64 cols, box_rows rows}; 128B swizzle; OOB -> 0.
int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return 0;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    MRA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand pointer must be 16-byte aligned");
    MRA_REQUIRE(ld % 8 == 0, "GEMM operand row stride must be a multiple of 8 elements (16 bytes), got %lld", (long long)ld);
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
                (long long)rows, (long long)cols, (long long)ld);
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 8192) cache.clear();
        cache[key] = m;
    }
    *out = m;
    return 0;
}

template <int BN, int STAGES, bool GELU, bool OUT_F32>
int launch_tc_variant(const GemmArgs& a, cudaStream_t s) {
    using L = SmemLayout<BN, STAGES>;
    auto kern = gemm_tc_kernel<BN, STAGES, GELU, OUT_F32>;
    static bool attr_set = false;
    if (!attr_set) {
        MRA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    if (int e = get_tensor_map(a.A, a.M, a.K, a.lda, BM, &tmA)) return e;
    if (int e = get_tensor_map(a.W, a.N, a.K, a.ldw, BN, &tmB)) return e;
    EpiParams p{a.bias, a.residual, a.ldr, a.C, a.ldc, a.M, a.N, a.K};
    const int m_tiles = (a.M + BM - 1) / BM, n_tiles = (a.N + BN - 1) / BN;
    const int total = m_tiles * n_tiles;
    const int grid = total < sm_count() ? total : sm_count();
    kern<<<grid, NUM_THREADS, L::TOTAL, s>>>(tmA, tmB, p);
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN, int STAGES>
int dispatch_epi(const GemmArgs& a, cudaStream_t s) {
    if (a.gelu) {
        if (a.out_fp32) return launch_tc_variant<BN, STAGES, true, true>(a, s);
        return launch_tc_variant<BN, STAGES, true, false>(a, s);
    }
    if (a.out_fp32) return launch_tc_variant<BN, STAGES, false, true>(a, s);
    return launch_tc_variant<BN, STAGES, false, false>(a, s);
}

int check_args(const GemmArgs& a) {
    MRA_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "GEMM with empty dimension M=%d N=%d K=%d", a.M, a.N, a.K);
    MRA_REQUIRE(a.N % 8 == 0, "GEMM N must be a multiple of 8, got %d", a.N);
    MRA_REQUIRE(a.K % 8 == 0, "GEMM K must be a multiple of 8, got %d", a.K);
    MRA_REQUIRE(a.ldc % 8 == 0, "GEMM ldc must be a multiple of 8, got %lld", (long long)a.ldc);
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(a.C) & 15) == 0, "GEMM output pointer must be 16-byte aligned");
    if (a.residual) {
        MRA_REQUIRE(a.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0,
                    "GEMM residual must be 16-byte aligned with ldr %% 4 == 0");
    }
    if (a.bias) MRA_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "GEMM bias must be 16-byte aligned");
    return 0;
}

}  // namespace

int launch_gemm_tc(const GemmArgs& a, cudaStream_t s) {
    if (int e = check_args(a)) return e;
    // Tile-width choice: 128x256 tiles run the tensor pipe at full rate (smem operand traffic 96 B/clk/SM); 128x128
    // tiles are smem-bound (128 B/clk) but quantise better when there are few tiles.  Pick the cheaper estimate.
    const int sms = sm_count();
    const long m_tiles = (a.M + BM - 1) / BM;
    const long t256 = m_tiles * ((a.N + 255) / 256), t128 = m_tiles * ((a.N + 127) / 128);
    const double c256 = double((t256 + sms - 1) / sms) * 2.0;
    const double c128 = double((t128 + sms - 1) / sms) * 1.15;
    if (a.N % 256 == 0 && c256 <= c128) return dispatch_epi<256, 4>(a, s);
    return dispatch_epi<128, 6>(a, s);
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t s) {
    if (int e = check_args(a)) return e;
    dim3 block(16, 16), grid((a.N + 15) / 16, (a.M + 15) / 16);
    EpiParams p{a.bias, a.residual, a.ldr, a.C, a.ldc, a.M, a.N, a.K};
    const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(a.A);
    const __nv_bfloat16* W = reinterpret_cast<const __nv_bfloat16*>(a.W);
    if (a.gelu) {
        if (a.out_fp32) gemm_simt_kernel<true, true><<<grid, block, 0, s>>>(A, a.lda, W, a.ldw, p);
        else gemm_simt_kernel<true, false><<<grid, block, 0, s>>>(A, a.lda, W, a.ldw, p);
    } else {
        if (a.out_fp32) gemm_simt_kernel<false, true><<<grid, block, 0, s>>>(A, a.lda, W, a.ldw, p);
        else gemm_simt_kernel<false, false><<<grid, block, 0, s>>>(A, a.lda, W, a.ldw, p);
    }
    MRA_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mra
