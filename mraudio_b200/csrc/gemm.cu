// C[M,N] = epilogue(A[M,K] . W[N,K]^T)  -- the Linear layers of the Q-Former and llm_proj.
//
// Blackwell-native design (sm_100a):
//   * persistent kernel, one CTA per SM, static round-robin over 128 x BN output tiles (n fastest so that the CTAs
//     running concurrently share the same A rows through L2);
//   * warp 0 = TMA producer (cp.async.bulk.tensor, 128B-swizzled 64-wide K slabs, STAGES-deep mbarrier ring),
//     warp 1 = MMA issuer (one thread, tcgen05.mma kind::f16, M=128, N=BN, K=16, fp32 accumulators in TMEM),
//     warps 2..9 = epilogue (two per TMEM lane quadrant, alternating over 128-byte-wide column chunks): tcgen05.ld
//     32x32b -> registers -> bias / erf-GELU / fp32 residual -> 128B-swizzled shared-memory staging -> TMA bulk tensor
//     STORE; the fp32 residual arrives by TMA LOAD into a per-warp chunk buffer, prefetched one chunk ahead (the
//     first one while the tile's main loop is still running).  All global traffic
//     of the epilogue is therefore asynchronous, fully coalesced and clipped at the M / N edges by the TMA unit;
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
// Reference arithmetic replaced: torch.nn.Linear (+ GELU, + residual add) in LAVIS Qformer.py / HF port
// modeling_instructblip.py:499-509,549-553,586-610 and llm_proj (models/xinstructblip.py:707-708).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.h"
#include "ptx.cuh"

// timing experiments (MRA_GEMM_DEBUG) are compiled in only with -DMRA_INSTRUMENT (make INSTRUMENT=1)
#ifdef MRA_INSTRUMENT
#define MRA_GEMM_DBG(p) ((p).dbg)
#else
#define MRA_GEMM_DBG(p) 0
#endif

namespace mra {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int CHUNK_BYTES = 32 * 128;       // one epilogue chunk of one warp: 32 rows x 128 bytes

constexpr int MAX_GROUPS = 4;

// A launch processes up to MAX_GROUPS independent problems that share N, K and the epilogue kind but have their own
// A / W / C / residual matrices, bias and M (grouped GEMM: e.g. FFN_query + FFN_text of both modalities in one launch,
// so that the persistent grid sees 4x the tiles and the wave-quantisation tail shrinks accordingly).
struct GroupMaps {
    CUtensorMap a[MAX_GROUPS], b[MAX_GROUPS], c[MAX_GROUPS], r[MAX_GROUPS];
};
struct EpiParams {
    const float* bias[MAX_GROUPS];
    int M[MAX_GROUPS];
    int c_frames[MAX_GROUPS];         // > 0: the C map of this group is the 4-D scatter map (GemmArgs::c_frames);
                                      // < 0: the 3-D head-major map (GemmArgs::c_head_major)
    int tile_start[MAX_GROUPS + 1];   // first tile index of each group (tiles of a group: m-block major, n fastest)
    int groups;
    int N, K;
    // ksplit > 1: the K range of every tile is cut into ksplit slices that are separate work items (fills the grid when
    // M x N gives few tiles and K is long: the weight gradients, K = number of tokens); needs reduce_add.
    // reduce_add: the epilogue adds its tile into C with TMA reduce-add stores (fp32) instead of storing it.
    int ksplit, kb_per_split, reduce_add;
    DropoutParams drop;               // thr8 != 0: dropout on (acc + bias) before the residual add (fp32 + residual epilogue)
    int drop_row0[MAX_GROUPS];        // first row of each group in the split token layout
    int dbg;   // -DMRA_INSTRUMENT + MRA_GEMM_DEBUG=8: cycles per tile spent waiting for the accumulator vs in the epilogue
};

__device__ unsigned long long g_gemm_timing[8];

__device__ __forceinline__ void decode_tile(const EpiParams& p, int tile, int n_tiles, int& g, int& m_blk, int& n_blk) {
    g = 0;
#pragma unroll
    for (int i = 1; i < MAX_GROUPS; ++i)
        if (i < p.groups && tile >= p.tile_start[i]) g = i;
    const int t = tile - p.tile_start[g];
    m_blk = t / n_tiles;
    n_blk = t - m_blk * n_tiles;
}

template <int BN, int STAGES, bool RES, bool U2 = false>
struct SmemLayout {
    static constexpr int B_STAGE_BYTES = (U2 ? BN / 2 : BN) * BK * 2;   // U2: each CTA of the pair keeps half of the W slab
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int EPI_BUFS_PER_WARP = RES ? 2 : 1;   // 1 output chunk (+ 1 residual chunk) of 4 KiB
    static constexpr int EPI_BYTES = EPI_WARPS * EPI_BUFS_PER_WARP * CHUNK_BYTES;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4 + EPI_WARPS;
    static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;  // + tmem ptr + 1024B alignment slack
    static_assert(TOTAL <= 227 * 1024, "shared memory budget exceeded");
};

// erf-GELU: 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the
// bf16 rounding of the result) -- 2 MUFU + ~12 FMA-pipe instructions per element instead of erff()'s ~30, which matters
// because the GELU epilogue is issue-bound (128 x BN elements per tile against a 6144-cycle main loop at BN = 256).
__device__ __forceinline__ float gelu_erf(float x) {
    const float u = x * 0.70710678118654752f;
    const float a = fabsf(u);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * a * -1.4426950408889634f));
    const float erf_abs = fmaf(-p, e, 1.0f);
    const float erf_u = copysignf(erf_abs, u);
    const float h = 0.5f * x;
    return fmaf(h, erf_u, h);
}

// Same function for bf16 outputs: 0.5 x (1 + erf(x / sqrt 2)) = x * (0.5 + 0.5 tanh(x (a + b x^2 + c x^4))) with (a, b, c)
// fitted to the erf form (|error| <= 2.6e-5 over all x; x^2 clamped at 64 where tanh has saturated, which also keeps
// the negative c from flipping the sign of the argument) and the hardware tanh.approx.f32 (relative error 2^-11):
// The total error (<= 5e-4 |x|) is an eighth of the bf16 rounding step of the stored result; fp32 outputs keep gelu_erf.
// Evaluated on a PAIR of values with packed fp32 instructions (FMUL2 / FFMA2): 6 FMA-pipe instructions + 2 MUFU + 2 FMNMX
// per pair instead of 14 + 2 + 2.
__device__ __forceinline__ ptx::f32x2_t gelu_fast2(ptx::f32x2_t x) {
    float u0, u1;
    ptx::unpack2(ptx::mul2(x, x), u0, u1);
    const ptx::f32x2_t u = ptx::pack2(fminf(u0, 64.0f), fminf(u1, 64.0f));
    ptx::f32x2_t q = ptx::fma2(ptx::pack2(-3.51534682e-4f, -3.51534682e-4f), u, ptx::pack2(3.70057307e-2f, 3.70057307e-2f));
    q = ptx::fma2(q, u, ptx::pack2(7.97507813e-1f, 7.97507813e-1f));
    float a0, a1, t0, t1;
    ptx::unpack2(ptx::mul2(x, q), a0, a1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
    return ptx::mul2(x, ptx::fma2(ptx::pack2(0.5f, 0.5f), ptx::pack2(t0, t1), ptx::pack2(0.5f, 0.5f)));
}

// byte offset of 16-byte unit `j` of row `r` inside a 32-row x 128-byte chunk buffer with the TMA 128-byte swizzle
__device__ __forceinline__ uint32_t swz(int r, int j) { return static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)); }

// CM = CTAs per cluster along M (1 or 2).  With CM = 2 the two CTAs of a cluster compute the tiles (2 mp, n) and
// (2 mp + 1, n): they need the same W rows, so each loads HALF of the W slab and multicasts it to both (the kernels are
// L2 -> SM bandwidth bound: 128 x 256 tiles need 96 B/clk/SM at full tensor rate; sharing W cuts that by a third).
// TN (operand form): 0 = A [M, K] and W [N, K], reduction index contiguous (the Linear forward);
//   1 = both operands given with the REDUCTION index as their row index (A: [K, M], W: [K, N], row-major): the
//       weight-gradient GEMM dW[out, in] = dY[n, out]^T . X[n, in] reads dY and X as they lie in memory;
//   2 = A [M, K] as in the forward, W [K, N] with the reduction index as its row index: the data-gradient GEMM
//       dX[n, in] = dY[n, out] . W[out, in] reads the forward's weight matrix as it lies.
// Operands whose reduction index is the row index go through MN-major shared-memory descriptors -- no transposed copies.
// U2 (with CM = 2): the pair of CTAs runs ONE tcgen05.mma.cta_group::2 with M = 256 per K step: each CTA holds its 128
// rows of A, HALF of the W slab (the tensor core reads the other half from the peer's shared memory) and its own
// 128 x BN accumulator.  A CTA then receives 32 KiB instead of 48 KiB per K slab for the same MMA work -- the plain
// kernel is bound by what an SM can take in (measured: the MMA thread waits for slabs arriving every ~620 cycles
// instead of 512, tools/gemm_phase.py) -- and the ring gets 6 stages instead of 4.  The leader (rank 0) issues the
// MMAs; both CTAs' TMA loads count their bytes on the leader's "full" barrier; tcgen05.commit multicasts the "slab
// free" and "accumulator ready" arrivals to both CTAs; the peer's epilogue warps release the accumulator remotely.
template <int BN, int STAGES, bool GELU, bool OUT_F32, bool RES, int CM, int TN = 0, bool U2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GroupMaps maps, const __grid_constant__ EpiParams p) {
    static_assert(!U2 || (CM == 2 && TN == 0), "the 2-CTA MMA form needs a pair of CTAs and K-major operands");
    using L = SmemLayout<BN, STAGES, RES, U2>;
    constexpr uint32_t TMEM_COLS = BN <= 128 ? 256 : 512;  // two accumulator stages, power of two
    constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;
    constexpr int CH = OUT_F32 ? 32 : 64;                  // columns per epilogue chunk (128 bytes of output per row)
    constexpr int NCH = BN / CH;
    static_assert(BN % CH == 0, "tile width must be a whole number of chunks");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_bar = tempty_bar + 2;                    // [EPI_WARPS]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tiles = (p.N + BN - 1) / BN;
    const int total_tiles = p.tile_start[p.groups];   // work items: (K slice, group, m-block [pair], n-block)
    const int total_items = total_tiles * p.ksplit;
    const uint32_t crank = CM > 1 ? ptx::cluster_ctarank() : 0u;
    const int first_item = CM > 1 ? static_cast<int>(blockIdx.x) / CM : static_cast<int>(blockIdx.x);
    const int item_stride = CM > 1 ? static_cast<int>(gridDim.x) / CM : static_cast<int>(gridDim.x);
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < p.groups; ++g) {
            ptx::prefetch_tensormap(&maps.a[g]);
            ptx::prefetch_tensormap(&maps.b[g]);
            ptx::prefetch_tensormap(&maps.c[g]);
            if (RES) ptx::prefetch_tensormap(&maps.r[g]);
        }
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], U2 ? 1 : CM);   // every CTA of the cluster must have consumed the slab
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], U2 ? 2 * EPI_WARPS : EPI_WARPS);  // one arrival per epilogue warp (of both CTAs)
        }
        for (int i = 0; i < EPI_WARPS; ++i) ptx::mbar_init(&res_bar[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        if (U2) {
            ptx::tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (CM > 1) ptx::cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above overlapped the tail of the previous kernel in the stream (programmatic dependent launch)
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = first_item; item < total_items; item += item_stride) {
                const int ks = item / total_tiles, tile = item - ks * total_tiles;
                int g, m_blk, n_blk;
                decode_tile(p, tile, n_tiles, g, m_blk, n_blk);
                m_blk = m_blk * CM + static_cast<int>(crank);
                const CUtensorMap* tmA = &maps.a[g];
                const CUtensorMap* tmB = &maps.b[g];
                const int kb0 = ks * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (U2) {
                        // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the pair
                        const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
                        if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
                        ptx::tma_load_2d_pair(sA + stage * A_STAGE_BYTES, tmA, lead_bar, kb * BK, m_blk * BM);
                        ptx::tma_load_2d_pair(sB + stage * L::B_STAGE_BYTES, tmB, lead_bar, kb * BK,
                                              n_blk * BN + static_cast<int>(crank) * (BN / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    if (TN != 0) {
                        // boxes of {64 MN elements, 64 reduction rows}: one per 64 output rows (A) / columns (W)
                        if (TN == 1) {
#pragma unroll
                            for (int i = 0; i < BM / 64; ++i)
                                ptx::tma_load_2d(sA + stage * A_STAGE_BYTES + i * 8192, tmA, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
                        } else {
                            ptx::tma_load_2d(sA + stage * A_STAGE_BYTES, tmA, &full_bar[stage], kb * BK, m_blk * BM);
                        }
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            ptx::tma_load_2d(sB + stage * L::B_STAGE_BYTES + j * 8192, tmB, &full_bar[stage], n_blk * BN + j * 64, kb * BK);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    ptx::tma_load_2d(sA + stage * A_STAGE_BYTES, tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    if (CM == 1) {
                        ptx::tma_load_2d(sB + stage * L::B_STAGE_BYTES, tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    } else {
                        // my half of the W slab, delivered to both CTAs of the pair (the box of tmB is BN / 2 rows)
                        ptx::tma_load_2d_multicast(sB + stage * L::B_STAGE_BYTES + crank * (L::B_STAGE_BYTES / 2), tmB,
                                                   &full_bar[stage], kb * BK, n_blk * BN + static_cast<int>(crank) * (BN / 2),
                                                   static_cast<uint16_t>(0x3));
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0 && (!U2 || crank == 0)) {
            constexpr uint32_t idesc = U2      ? ptx::make_idesc_bf16_f32(2 * BM, BN)
                                       : TN == 1 ? ptx::make_idesc_bf16_f32_mn(BM, BN)
                                       : TN == 2 ? (ptx::make_idesc_bf16_f32(BM, BN) | (1u << 16))   // b_major = MN only
                                                 : ptx::make_idesc_bf16_f32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int item = first_item; item < total_items; item += item_stride, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                const bool mprof = (MRA_GEMM_DBG(p) & 8) && blockIdx.x == 0;
                long long mt0 = 0, mt1 = 0, mt2 = 0;
                if (mprof) mt0 = clock64();
                ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                ptx::tc_fence_after();
                if (mprof) mt1 = clock64();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                const int kb0 = (item / total_tiles) * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    if (mprof && kb == kb0) mt2 = clock64();
                    ptx::tc_fence_after();
                    if (TN != 0) {
                        const uint64_t a_desc = TN == 1 ? ptx::make_sw128_mnmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES), 8192)
                                                        : ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
                        const uint64_t b_desc = ptx::make_sw128_mnmajor_desc(ptx::smem_u32(sB + stage * L::B_STAGE_BYTES), 8192);
                        constexpr int A_STEP = TN == 1 ? 128 : 2;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // MN-major: advance 16 reduction rows of 128 B: +2048 B = +128 in the >>4 address field
                            ptx::umma_bf16_ss(d_tmem, a_desc + A_STEP * k, b_desc + 128 * k, idesc, kb > kb0 || k != 0);
                        }
                    } else {
                        const uint64_t a_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
                        const uint64_t b_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + stage * L::B_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the >>4 address field
                            if (U2) ptx::umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, kb > kb0 || k != 0);
                            else ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, kb > kb0 || k != 0);
                        }
                    }
                    // smem slot reusable once these MMAs have read it (CM = 2: the peer multicasts into it too)
                    if (U2) ptx::umma_commit_pair(&empty_bar[stage], static_cast<uint16_t>(0x3));
                    else if (CM == 1) ptx::umma_commit(&empty_bar[stage]);
                    else ptx::umma_commit_multicast(&empty_bar[stage], static_cast<uint16_t>(0x3));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (U2: in both CTAs of the pair)
                if (U2) ptx::umma_commit_pair(&tfull_bar[acc], static_cast<uint16_t>(0x3));
                else ptx::umma_commit(&tfull_bar[acc]);
                if (mprof) {
                    const long long mt3 = clock64();
                    g_gemm_timing[3] += mt1 - mt0; g_gemm_timing[4] += mt2 - mt1; g_gemm_timing[5] += mt3 - mt2;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..9
        // Two warps share each TMEM lane quadrant (32 rows) and alternate over the tile's column chunks, so that every
        // SM sub-partition has two epilogue warps to hide TMEM / shared-memory / MUFU latencies behind each other.
        const int quad = warp & 3;      // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
        const int ew = warp - 2;        // private staging buffers / residual barrier
        const int member = ew >> 2;     // 0 / 1: takes the chunks c == member (mod 2)
        uint8_t* my = smem + L::EPI_OFFSET + ew * L::EPI_BUFS_PER_WARP * CHUNK_BYTES;   // [out]([residual] when RES)
        uint8_t* odst = my;
        uint8_t* rsrc = my + CHUNK_BYTES;
        const uint32_t odst_s = ptx::smem_u32(odst), rsrc_s = ptx::smem_u32(rsrc);
        uint64_t* rbar = res_bar + ew;
        uint32_t rphase = 0;
        constexpr int RSUB = CH / 32;   // 32-column residual sub-chunks per output chunk
        int iter = 0;
        for (int item = first_item; item < total_items; item += item_stride, ++iter) {
            const int tile = item % total_tiles;
            int g, m_blk, n_blk;
            decode_tile(p, tile, n_tiles, g, m_blk, n_blk);
            m_blk = m_blk * CM + static_cast<int>(crank);
            const CUtensorMap* tmC = &maps.c[g];
            const CUtensorMap* tmR = &maps.r[g];
            const float* bias = p.bias[g];
            const int Mg = p.M[g];
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int row0 = m_blk * BM + quad * 32;   // first row of this warp's slab
            const int col_base = n_blk * BN;
            if (RES && lane == 0 && member < NCH) {
                // first residual sub-chunk of this tile: issued before the accumulator is ready (overlaps the main loop)
                ptx::mbar_arrive_expect_tx(rbar, CHUNK_BYTES);
                ptx::tma_load_2d(rsrc, tmR, rbar, col_base + member * CH, row0);
            }
            const bool prof = (MRA_GEMM_DBG(p) & 8) && blockIdx.x == 0 && ew == 0 && lane == 0;
            long long pt0 = 0, pt1 = 0;
            if (prof) pt0 = clock64();
            // Bias of this warp's chunks, fetched BEFORE the wait for the accumulator (the loads used to sit on the critical
            // path of every chunk: ~1 k cycles per tile, tools/gemm_phase.py): lane l keeps columns l * BPL .. of each chunk
            // and the values are handed out by shuffles below.
            constexpr int MAXC = (NCH + 1) / 2;     // chunks of a tile per warp
            constexpr int BPL = CH / 32;            // bias values per lane and chunk
            float bq[MAXC][BPL];
#pragma unroll
            for (int ci = 0; ci < MAXC; ++ci) {
                const int c = member + 2 * ci;
#pragma unroll
                for (int e = 0; e < BPL; ++e) {
                    const int col = col_base + c * CH + lane * BPL + e;
                    bq[ci][e] = (bias != nullptr && c < NCH && col < p.N) ? __ldg(bias + col) : 0.f;
                }
            }
            ptx::mbar_wait(&tfull_bar[acc], acc_phase);
            if (prof) pt1 = clock64();
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_STRIDE;
#pragma unroll
            for (int ci = 0; ci < MAXC; ++ci) {
                const int c = member + 2 * ci;
                if (c >= NCH) break;
                const int col0 = col_base + c * CH;
                // all TMEM reads of the chunk are issued before the single wait (one exposed TMEM latency per chunk, not per half)
                uint32_t racc[RSUB][32];
#pragma unroll
                for (int half = 0; half < RSUB; ++half) {
                    if (MRA_GEMM_DBG(p) & 16) {   // experiment: no TMEM reads (results are garbage) -- isolates the drain's cost
#pragma unroll
                        for (int j = 0; j < 32; ++j) racc[half][j] = 0;
                    } else {
                        ptx::tmem_ld_32x32b_x32(t_row + c * CH + half * 32, racc[half]);
                    }
                }
                if (!(MRA_GEMM_DBG(p) & 16)) ptx::tmem_ld_wait();
#pragma unroll
                for (int half = 0; half < RSUB; ++half) {
                    float v[32];
                    if constexpr (!(GELU && !OUT_F32)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(racc[half][j]);
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int idx = half * 32 + j;      // column inside the chunk -> (lane, slot) that holds its bias
                            v[j] += __shfl_sync(0xffffffffu, bq[ci][idx % BPL], idx / BPL);
                        }
                    }
                    if (GELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                    }
                    } else {
                    // bf16 GELU epilogue (FFN-up): pairs of columns through the packed fp32 pipe (FADD2 / FMUL2 / FFMA2) -- this
                    // epilogue is issue-bound on the FMA pipe: 64.5 vs 67.6 us on the FFN-up shape of a step, same box.  (The
                    // epilogues without GELU are not: packing them changes nothing.)  The bias test is hoisted out of the loop
                    // so that the pairs' dependency chains interleave freely.
                    ptx::f32x2_t xx[16];
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const int idx = half * 32 + j;      // column inside the chunk -> (lane, slot) that holds its bias
                            xx[j >> 1] = ptx::add2(ptx::pack2(__uint_as_float(racc[half][j]), __uint_as_float(racc[half][j + 1])),
                                                   ptx::pack2(__shfl_sync(0xffffffffu, bq[ci][idx % BPL], idx / BPL),
                                                              __shfl_sync(0xffffffffu, bq[ci][(idx + 1) % BPL], (idx + 1) / BPL)));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 2)
                            xx[j >> 1] = ptx::pack2(__uint_as_float(racc[half][j]), __uint_as_float(racc[half][j + 1]));
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) ptx::unpack2(gelu_fast2(xx[j >> 1]), v[j], v[j + 1]);
                    }
                    if constexpr (RES && OUT_F32) {
                        if (p.drop.thr8 != 0) {
                            // training-mode dropout of BertSelfOutput / BertOutput: on the Linear's output, before the residual
                            const uint64_t grow = static_cast<uint64_t>(p.drop_row0[g] + row0 + lane);
                            const uint64_t base = grow * static_cast<uint64_t>(p.N >> 3) + ((col0 + half * 32) >> 3);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 rb = dropout_bytes(p.drop, base + q);
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[8 * q + e] *= dropout_mult(p.drop, rb, e);
                            }
                        }
                    }
                    if (RES) {
                        ptx::mbar_wait(rbar, rphase);
                        rphase ^= 1;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 x = ptx::ld_shared_v4f(rsrc_s + swz(lane, j));
                            v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
                        }
                        __syncwarp();   // every lane has read the residual buffer: refill it with the next sub-chunk
                        if (lane == 0) {
                            int nc = c, nh = half + 1;
                            if (nh == RSUB) { nh = 0; nc += 2; }
                            if (nc < NCH) {
                                ptx::mbar_arrive_expect_tx(rbar, CHUNK_BYTES);
                                ptx::tma_load_2d(rsrc, tmR, rbar, col_base + nc * CH + nh * 32, row0);
                            }
                        }
                    }
                    if (OUT_F32 || half == 0) {
                        // the TMA store that last read this warp's staging buffer must have finished reading shared memory
                        // (waited for here, after the TMEM read and the arithmetic of this chunk, not before them)
                        if (lane == 0) ptx::tma_store_wait_read<0>();
                        __syncwarp();
                    }
                    if (OUT_F32) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            ptx::st_shared_v4f(odst_s + swz(lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ptx::st_shared_v4(odst_s + swz(lane, half * 4 + j), ptx::pack_bf16x2(v[8 * j], v[8 * j + 1]),
                                              ptx::pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), ptx::pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                                              ptx::pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                        }
                    }
                }
                if (c + 2 >= NCH) {
                    // this warp's tcgen05.ld of the tile are done: hand the accumulator back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (U2) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[acc]), 0));   // the leader's barrier
                        else ptx::mbar_arrive(&tempty_bar[acc]);
                    }
                }
                ptx::fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (col0 < p.N && row0 < Mg) {
                        bool scattered = false;
                        if constexpr (!OUT_F32) {
                            const int cf = p.c_frames[g];
                            if (cf > 0) {   // 32-row slab = the 32 query tokens of one (video, frame): coordinates (col, q, f, b)
                                const int rf = row0 >> 5;
                                ptx::tma_store_4d(tmC, odst, col0, 0, rf % cf, rf / cf);
                                scattered = true;
                            } else if (cf < 0) {   // head-major C: this 64-column chunk is 32 full rows of slot col0 / 64
                                ptx::tma_store_3d(tmC, odst, 0, row0, col0 >> 6);
                                scattered = true;
                            }
                        }
                        if (!scattered) {
                            if (OUT_F32 && p.reduce_add) ptx::tma_reduce_add_2d(tmC, odst, col0, row0);
                            else ptx::tma_store_2d(tmC, odst, col0, row0);  // edges clipped by TMA
                        }
                    }
                    ptx::tma_store_commit();
                }
            }
            if (prof) {
                const long long pt2 = clock64();
                g_gemm_timing[0] += pt1 - pt0; g_gemm_timing[1] += pt2 - pt1; g_gemm_timing[2] += 1;
            }
            if (member >= NCH) {   // (only when a tile has a single chunk) nothing to read: release immediately
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                        if (U2) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[acc]), 0));   // the leader's barrier
                        else ptx::mbar_arrive(&tempty_bar[acc]);
                    }
            }
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (CM > 1) ptx::cluster_sync_all();   // no CTA exits while its peer may still multicast into it
    if (warp == 1) {
        ptx::tc_fence_after();
        if (U2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Debug / test-only CUDA-core kernel with the same contract (isolates descriptor bugs in the tensor-core path).
struct SimtParams {
    const float* bias;
    const float* residual;
    int64_t ldr;
    void* C;
    int64_t ldc;
    int M, N, K;
};

template <bool GELU, bool OUT_F32>
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda, const __nv_bfloat16* __restrict__ W,
                                 int64_t ldw, const SimtParams p) {
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
        if (GELU) acc = 0.5f * acc * (1.0f + erff(acc * 0.70710678118654752f));  // library erf: independent of gelu_erf
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
    int box_rows, box_cols, elem_bytes;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
               box_cols == o.box_cols && elem_bytes == o.elem_bytes;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        h = h * 1000003u ^ static_cast<size_t>(k.cols);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.box_rows * 131 + k.box_cols * 7 + k.elem_bytes);
        return h;
    }
};

// Row-major [rows, cols] matrix of 2-byte (bf16) or 4-byte (fp32) elements with row stride ld (elements);
// box = {box_cols, box_rows} with box_cols * elem_bytes == 128; 128B swizzle; out-of-bounds reads -> 0, writes clipped.
}  // namespace

int get_tensor_map(const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, int elem_bytes,
                   CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, rows, cols, ld, box_rows, box_cols, elem_bytes};
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
    MRA_REQUIRE((ld * elem_bytes) % 16 == 0, "GEMM operand row stride must be a multiple of 16 bytes, got %lld elements of %d bytes",
                (long long)ld, elem_bytes);
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

int get_tensor_map_scatter(const void* ptr, int64_t batch, int64_t frames, int64_t cols, int64_t ld, int64_t frame_stride,
                           int64_t batch_stride, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode_fn();
    MRA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && frame_stride % 8 == 0 && batch_stride % 8 == 0,
                "scatter output: pointer and strides must be multiples of 16 bytes");
    MRA_REQUIRE(batch > 0 && frames > 0 && ld >= cols && frame_stride >= 32 * ld && (batch == 1 || batch_stride >= frames * frame_stride),
                "scatter output: overlapping slots (ld=%lld frame_stride=%lld batch_stride=%lld)", (long long)ld,
                (long long)frame_stride, (long long)batch_stride);
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), 32, static_cast<cuuint64_t>(frames), static_cast<cuuint64_t>(batch)};
    cuuint64_t gstride[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(frame_stride) * 2,
                             static_cast<cuuint64_t>(batch > 1 ? batch_stride : frames * frame_stride) * 2};
    cuuint32_t box[4] = {64, 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D scatter) failed with CUresult %d", (int)r);
    return 0;
}

// bf16 [cols / 64][rows][64]: 64-column slot c / 64 of output row r at ptr + ((c / 64) * rows + r) * 64; box {64, 32, 1}
int get_tensor_map_head_major(const void* ptr, int64_t rows, int64_t cols, CUtensorMap* out) {
    EncodeTiledFn enc = get_encode_fn();
    MRA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && rows > 0 && cols > 0 && cols % 64 == 0,
                "head-major output: pointer must be 16-byte aligned, cols a multiple of 64");
    cuuint64_t gdim[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 64)};
    cuuint64_t gstride[2] = {128, static_cast<cuuint64_t>(rows) * 128};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MRA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (head-major output) failed with CUresult %d", (int)r);
    return 0;
}

namespace {

template <int BN, int STAGES, bool GELU, bool OUT_F32, bool RES, int CM, int TN = 0, bool U2 = false>
int launch_tc_variant(const GemmArgs* ga, int n, cudaStream_t s) {
    using L = SmemLayout<BN, STAGES, RES, U2>;
    auto kern = gemm_tc_kernel<BN, STAGES, GELU, OUT_F32, RES, CM, TN, U2>;
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(kern), L::TOTAL)) return e;
    GroupMaps maps;
    EpiParams p;
    p.groups = n;
    p.N = ga[0].N;
    p.K = ga[0].K;
    p.reduce_add = ga[0].reduce_add;
    p.drop = ga[0].drop;
    MRA_REQUIRE(p.drop.thr8 == 0 || (RES && OUT_F32 && !GELU && TN == 0), "GEMM dropout needs the fp32 + residual epilogue");
    static const int dbg_env = [] { const char* e = getenv("MRA_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg_env;
    {
        const int k_blocks = (p.K + BK - 1) / BK;
        int ks = ga[0].ksplit < 1 ? 1 : ga[0].ksplit;
        if (ks > k_blocks) ks = k_blocks;
        p.kb_per_split = (k_blocks + ks - 1) / ks;
        p.ksplit = (k_blocks + p.kb_per_split - 1) / p.kb_per_split;   // no empty slice
        MRA_REQUIRE(p.ksplit == 1 || (p.reduce_add && OUT_F32 && !RES && !GELU),
                    "split-K needs the fp32 reduce-add epilogue (no residual, no GELU)");
    }
    const int n_tiles = (p.N + BN - 1) / BN;
    int total = 0;
    for (int g = 0; g < MAX_GROUPS; ++g) {
        const GemmArgs& a = ga[g < n ? g : 0];
        if (g < n) {
            if (TN != 0) {
                if (TN == 1) {
                    if (int e = get_tensor_map(a.A, a.K, a.M, a.lda, BK, 64, 2, &maps.a[g])) return e;
                } else if (int e = get_tensor_map(a.A, a.M, a.K, a.lda, BM, BK, 2, &maps.a[g])) return e;
                if (int e = get_tensor_map(a.W, a.K, a.N, a.ldw, BK, 64, 2, &maps.b[g])) return e;
            } else {
                if (int e = get_tensor_map(a.A, a.M, a.K, a.lda, BM, BK, 2, &maps.a[g])) return e;
                if (int e = get_tensor_map(a.W, a.N, a.K, a.ldw, BN / CM, BK, 2, &maps.b[g])) return e;
            }
            p.c_frames[g] = a.c_head_major ? -1 : a.c_frames;
            if (a.c_head_major) {
                MRA_REQUIRE(!OUT_F32 && a.c_frames == 0 && a.N % 64 == 0, "head-major output needs bf16 C and N a multiple of 64 (N=%d)", a.N);
                if (int e = get_tensor_map_head_major(a.C, a.M, a.N, &maps.c[g])) return e;
            } else if (a.c_frames > 0) {
                MRA_REQUIRE(!OUT_F32 && a.M % 32 == 0 && (a.M / 32) % a.c_frames == 0,
                            "scatter output needs bf16 C and M = videos * frames * 32 (M=%d frames=%d)", a.M, a.c_frames);
                if (int e = get_tensor_map_scatter(a.C, a.M / 32 / a.c_frames, a.c_frames, a.N, a.ldc, a.c_frame_stride,
                                                   a.c_batch_stride, &maps.c[g])) return e;
            } else if (int e = get_tensor_map(a.C, a.M, a.N, a.ldc, 32, OUT_F32 ? 32 : 64, OUT_F32 ? 4 : 2, &maps.c[g])) return e;
            if (RES) {
                if (int e = get_tensor_map(a.residual, a.M, a.N, a.ldr, 32, 32, 4, &maps.r[g])) return e;
            } else {
                maps.r[g] = maps.c[g];
            }
            p.bias[g] = a.bias;
            p.M[g] = a.M;
            p.drop_row0[g] = a.drop_row0;
            p.tile_start[g] = total;
            total += (((a.M + BM - 1) / BM + CM - 1) / CM) * n_tiles;   // work items: m-block pairs when CM == 2
        } else {
            maps.a[g] = maps.a[0]; maps.b[g] = maps.b[0]; maps.c[g] = maps.c[0]; maps.r[g] = maps.r[0];
            p.bias[g] = nullptr;
            p.M[g] = 0;
            p.drop_row0[g] = 0;
            p.c_frames[g] = 0;
            p.tile_start[g] = total;
        }
    }
    p.tile_start[MAX_GROUPS] = total;
    for (int g = n; g <= MAX_GROUPS; ++g) p.tile_start[g] = total;
    if (CM == 1) {
        const long items = static_cast<long>(total) * p.ksplit;
        const int grid = items < sm_count() ? static_cast<int>(items) : sm_count();
        MRA_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), L::TOTAL, s, 1, maps, p));
        if (MRA_GEMM_DBG(p) & 8) {   // (single-CTA path; the paired paths print below)
            unsigned long long t[8];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(t, g_gemm_timing, sizeof(t));
            const double nt = t[2] ? double(t[2]) : 1.0;
            fprintf(stderr, "[gemm timing BN=%d M=%d N=%d K=%d: %llu tiles on CTA 0] epilogue warp: wait-accumulator %.0f | work %.0f ; "
                            "MMA thread: wait-tmem-free %.0f | wait-first-slab %.0f | issue-loop %.0f cycles per tile\n",
                    BN, ga[0].M, p.N, p.K, t[2], t[0] / nt, t[1] / nt, t[3] / nt, t[4] / nt, t[5] / nt);
            unsigned long long z[8] = {0};
            cudaMemcpyToSymbol(g_gemm_timing, z, sizeof(z));
        }
        return 0;
    }
    const int clusters = total < sm_count() / CM ? total : sm_count() / CM;
    MRA_CHECK_CUDA(launch_pdl(kern, dim3(clusters * CM), dim3(NUM_THREADS), L::TOTAL, s, CM, maps, p));
    if (MRA_GEMM_DBG(p) & 8) {
        unsigned long long t[8];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(t, g_gemm_timing, sizeof(t));
        const double nt = t[2] ? double(t[2]) : 1.0;
        fprintf(stderr, "[gemm timing BN=%d CM=%d U2=%d M=%d N=%d K=%d: %llu tiles on CTA 0] epilogue warp: wait-accumulator %.0f | work %.0f ; "
                        "MMA thread: wait-tmem-free %.0f | wait-first-slab %.0f | issue-loop %.0f cycles per tile\n",
                BN, CM, int(U2), ga[0].M, p.N, p.K, t[2], t[0] / nt, t[1] / nt, t[3] / nt, t[4] / nt, t[5] / nt);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_gemm_timing, z, sizeof(z));
    }
    return 0;
}

template <int BN, int ST_PLAIN, int ST_RES, int CM, bool U2 = false>
int dispatch_epi(const GemmArgs* a, int n, cudaStream_t s) {
    const int code = (a[0].gelu ? 4 : 0) | (a[0].out_fp32 ? 2 : 0) | (a[0].residual ? 1 : 0);
    switch (code) {
        case 0: return launch_tc_variant<BN, ST_PLAIN, false, false, false, CM, 0, U2>(a, n, s);
        case 1: return launch_tc_variant<BN, ST_RES, false, false, true, CM, 0, U2>(a, n, s);
        case 2: return launch_tc_variant<BN, ST_PLAIN, false, true, false, CM, 0, U2>(a, n, s);
        case 3: return launch_tc_variant<BN, ST_RES, false, true, true, CM, 0, U2>(a, n, s);
        case 4: return launch_tc_variant<BN, ST_PLAIN, true, false, false, CM, 0, U2>(a, n, s);
        case 5: return launch_tc_variant<BN, ST_RES, true, false, true, CM, 0, U2>(a, n, s);
        case 6: return launch_tc_variant<BN, ST_PLAIN, true, true, false, CM, 0, U2>(a, n, s);
        default: return launch_tc_variant<BN, ST_RES, true, true, true, CM, 0, U2>(a, n, s);
    }
}

int check_args(const GemmArgs& a) {
    MRA_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "GEMM with empty dimension M=%d N=%d K=%d", a.M, a.N, a.K);
    MRA_REQUIRE(a.N % 8 == 0, "GEMM N must be a multiple of 8, got %d", a.N);
    MRA_REQUIRE(a.tn >= 0 && a.tn <= 2, "GEMM operand form must be 0, 1 or 2, got %d", a.tn);
    MRA_REQUIRE(a.tn == 1 || a.K % 8 == 0, "GEMM K must be a multiple of 8, got %d", a.K);
    MRA_REQUIRE(a.tn != 1 || a.M % 8 == 0, "transposed-operand GEMM: M must be a multiple of 8, got %d", a.M);
    MRA_REQUIRE((a.ldc * (a.out_fp32 ? 4 : 2)) % 16 == 0, "GEMM output row stride must be a multiple of 16 bytes, got ldc=%lld",
                (long long)a.ldc);
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(a.C) & 15) == 0, "GEMM output pointer must be 16-byte aligned");
    if (a.residual) {
        MRA_REQUIRE(a.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0,
                    "GEMM residual must be 16-byte aligned with ldr %% 4 == 0");
    }
    if (a.bias) MRA_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "GEMM bias must be 16-byte aligned");
    return 0;
}

}  // namespace

static int g_forced_bn = [] { const char* e = getenv("MRA_GEMM_BN"); return e ? atoi(e) : 0; }();
// 3 (default) = pairs of CTAs running tcgen05.mma.cta_group::2; 2 = pairs sharing the W slab by multicast; 1 = single CTAs
static int g_cluster_m = [] { const char* e = getenv("MRA_GEMM_CLUSTER"); return e ? atoi(e) : 3; }();
void set_gemm_tile_override(int bn) { g_forced_bn = bn; }
void set_gemm_cluster_override(int cm) { g_cluster_m = cm; }

int launch_gemm_tc_grouped(const GemmArgs* a, int n, cudaStream_t s) {
    MRA_REQUIRE(n >= 1 && n <= MAX_GROUPS, "grouped GEMM takes 1..%d problems, got %d", MAX_GROUPS, n);
    for (int g = 0; g < n; ++g) {
        if (int e = check_args(a[g])) return e;
        MRA_REQUIRE(a[g].N == a[0].N && a[g].K == a[0].K && a[g].gelu == a[0].gelu && a[g].out_fp32 == a[0].out_fp32 &&
                        (a[g].residual != nullptr) == (a[0].residual != nullptr) && a[g].tn == a[0].tn,
                    "grouped GEMM problems must share N, K and the epilogue kind");
    }
    // Tile-width choice by a wave-quantisation estimate: cost = waves x (tile width) x (a factor for how well that
    // width feeds the tensor pipe: 128-wide tiles are shared-memory-bandwidth bound).  MRA_GEMM_BN overrides (tuning).
    const int sms = sm_count();
    long m_tiles = 0;
    for (int g = 0; g < n; ++g) m_tiles += (a[g].M + BM - 1) / BM;
    const int forced = g_forced_bn;
    int best_bn = 128;
    double best = 1e30;
    const int cand[3] = {256, 192, 128};
    const double factor[3] = {1.0, 1.05, 1.15};
    for (int i = 0; i < 3; ++i) {
        const int bn = cand[i];
        if (forced ? bn != forced : false) continue;
        const long tiles = m_tiles * ((a[0].N + bn - 1) / bn);
        const double cost = double((tiles + sms - 1) / sms) * bn * factor[i];
        if (cost < best) { best = cost; best_bn = bn; }
    }
    if (a[0].tn == 1) {
        for (int g = 0; g < n; ++g)
            MRA_REQUIRE(a[g].out_fp32 && !a[g].gelu, "transposed-operand GEMM (tn = 1) needs fp32 output and no GELU");
        bool res = a[0].residual != nullptr;
        // "dW += dY^T X" (residual == C): accumulate with TMA reduce-add stores instead of reading C back, and cut the
        // token dimension into slices so that the few (out x in) tiles of a weight matrix fill the 148 SMs
        bool in_place = res;
        for (int g = 0; g < n; ++g) in_place = in_place && a[g].residual == a[g].C && a[g].ldr == a[g].ldc && a[g].bias == nullptr;
        static const bool no_split = getenv("MRA_NO_SPLITK") != nullptr;
        GemmArgs split[MAX_GROUPS];
        if (in_place && !no_split) {
            const long tiles = m_tiles * ((a[0].N + best_bn - 1) / best_bn);
            const int k_blocks = (a[0].K + BK - 1) / BK;
            int ks = static_cast<int>(sms / (tiles > 0 ? tiles : 1));
            if (ks > k_blocks / 4) ks = k_blocks / 4;     // at least 4 K slabs per slice
            if (ks < 1) ks = 1;
            for (int g = 0; g < n; ++g) {
                split[g] = a[g];
                split[g].residual = nullptr;
                split[g].ksplit = ks;
                split[g].reduce_add = 1;
            }
            a = split;
            res = false;
        }
        if (best_bn == 256) return res ? launch_tc_variant<256, 3, false, true, true, 1, 1>(a, n, s)
                                       : launch_tc_variant<256, 4, false, true, false, 1, 1>(a, n, s);
        if (best_bn == 192) return res ? launch_tc_variant<192, 4, false, true, true, 1, 1>(a, n, s)
                                       : launch_tc_variant<192, 4, false, true, false, 1, 1>(a, n, s);
        return res ? launch_tc_variant<128, 5, false, true, true, 1, 1>(a, n, s)
                   : launch_tc_variant<128, 6, false, true, false, 1, 1>(a, n, s);
    }
    if (a[0].tn == 2) {
        const bool res = a[0].residual != nullptr, f32 = a[0].out_fp32 != 0;
        MRA_REQUIRE(!a[0].gelu && (f32 || !res), "data-gradient GEMM (tn = 2): no GELU, residual only with fp32 output");
        if (best_bn == 256)
            return !f32 ? launch_tc_variant<256, 4, false, false, false, 1, 2>(a, n, s)
                        : res ? launch_tc_variant<256, 3, false, true, true, 1, 2>(a, n, s)
                              : launch_tc_variant<256, 4, false, true, false, 1, 2>(a, n, s);
        if (best_bn == 192)
            return !f32 ? launch_tc_variant<192, 4, false, false, false, 1, 2>(a, n, s)
                        : res ? launch_tc_variant<192, 4, false, true, true, 1, 2>(a, n, s)
                              : launch_tc_variant<192, 4, false, true, false, 1, 2>(a, n, s);
        return !f32 ? launch_tc_variant<128, 6, false, false, false, 1, 2>(a, n, s)
                    : res ? launch_tc_variant<128, 5, false, true, true, 1, 2>(a, n, s)
                          : launch_tc_variant<128, 6, false, true, false, 1, 2>(a, n, s);
    }
    // Pairs of CTAs along M when every problem has enough row blocks.  Measured on B200: sharing the W slab by multicast
    // (mode 2) neither helps nor hurts -- each SM still takes in 48 KiB per K slab -- while the 2-CTA MMA (mode 3), where
    // a CTA takes in 32 KiB, lifts the cross-K/V GEMM from 1412 to 1542 TF/s and the FFN-down shape from 868 to 1063.
    bool pair = g_cluster_m != 1;
    for (int g = 0; g < n; ++g) pair = pair && a[g].M >= 4 * BM;
    if (pair && g_cluster_m == 3) {
        // 2-CTA MMA (cta_group::2): half of the W slab per CTA, 6-stage ring
        if (best_bn == 256) return dispatch_epi<256, 6, 5, 2, true>(a, n, s);
        if (best_bn == 192) return dispatch_epi<192, 6, 5, 2, true>(a, n, s);
        return dispatch_epi<128, 6, 6, 2, true>(a, n, s);
    }
    if (pair) {
        if (best_bn == 256) return dispatch_epi<256, 4, 3, 2>(a, n, s);
        if (best_bn == 192) return dispatch_epi<192, 4, 4, 2>(a, n, s);
        return dispatch_epi<128, 6, 5, 2>(a, n, s);
    }
    if (best_bn == 256) return dispatch_epi<256, 4, 3, 1>(a, n, s);
    if (best_bn == 192) return dispatch_epi<192, 4, 4, 1>(a, n, s);
    return dispatch_epi<128, 6, 5, 1>(a, n, s);
}

int launch_gemm_tc(const GemmArgs& a, cudaStream_t s) { return launch_gemm_tc_grouped(&a, 1, s); }

int launch_gemm_simt(const GemmArgs& a, cudaStream_t s) {
    if (int e = check_args(a)) return e;
    dim3 block(16, 16), grid((a.N + 15) / 16, (a.M + 15) / 16);
    SimtParams p{a.bias, a.residual, a.ldr, a.C, a.ldc, a.M, a.N, a.K};
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
