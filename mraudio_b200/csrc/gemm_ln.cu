// Y = LayerNorm(A . W^T + bias + residual) * gamma + beta   for N = 768 (the Q-Former hidden size) in ONE kernel:
// the out-projection / FFN-down Linear of every Q-Former block together with its residual add and post-LayerNorm
// (BertSelfOutput / BertOutput, HF port modeling_instructblip.py:549-553, 606-610).  The pre-LayerNorm sums never
// touch HBM: they go from TMEM to registers, are normalised there and leave as the fp32 + bf16 LayerNorm outputs.
//
// A 128 x 768 fp32 accumulator needs 768 TMEM columns (an SM has 512).  A thread-block CLUSTER of THREE CTAs owns one
// 128-row block: CTA r computes columns [256 r, 256 r + 256) -- one tcgen05.mma (M=128, N=256) per 16-wide K step, TWO
// accumulator stages in TMEM (2 x 256 columns), so the epilogue of block i overlaps the main loop of block i+1 exactly
// as in gemm.cu -- and the per-row LayerNorm statistics are combined across the three CTAs through distributed shared
// memory (st.async ... mbarrier::complete_tx into the peers' stats buffers).  Per CTA:
//   warp 0      TMA producer: A tile 128 x 64 + W tile 256 x 64 per stage (48 KiB), 3 stages
//   warp 1      MMA issuer (single thread)
//   warps 4..11 epilogue (two warpgroups, 232 registers per thread via setmaxnreg), thread = (row, 128-column segment):
//       pass A  tcgen05.ld (once; TMEM reads cost 64 B/clk) -> + bias + residual (fp32, TMA-loaded 32-column chunks);
//               the 128 values stay in registers; row sum / sum of squares; accumulator handed back to the MMA warp
//       exchange partial statistics inside the CTA and with the two peer CTAs
//       pass B  normalise from registers, * gamma + beta -> fp32 and bf16 copies through swizzled staging + TMA stores
// (An earlier 2-CTA version with 384 columns per CTA could not double-buffer TMEM and serialised main loop and
//  epilogue: 92 us on the AO shape against 2 x 64 us unfused; see profiles/.)  Up to 4 grouped problems per launch.
//
// SPLIT = true: the residual stream is a PAIR of bf16 tensors (hi = bf16(y), lo = bf16(y - hi); hi + lo carries 16 mantissa
// bits) instead of fp32 + a bf16 shadow: hi is at the same time the next GEMM's A operand, so a post-LayerNorm element costs
// 4 B read + 4 B written instead of 4 + 6, the epilogue moves 64-column chunks (one 128-byte line per row and tensor) and
// pass B needs one staging turnaround per block instead of three.  This is the form the inference forward uses.
#include "common.h"
#include "ptx.cuh"

// timing experiments (MRA_LN_DEBUG) are compiled in only with -DMRA_INSTRUMENT (make INSTRUMENT=1)
#ifdef MRA_INSTRUMENT
#define MRA_LN_DBG(p) ((p).dbg)
#else
#define MRA_LN_DBG(p) 0
#endif

namespace mra {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NTOT = 768;                   // LayerNorm width handled by a cluster
constexpr int NCTA = 256;                   // columns per CTA (= one MMA)
constexpr int NSEG = 128;                   // columns per epilogue thread
constexpr int NSEGS = NTOT / NSEG;          // 6 partial-statistics sources per row
constexpr int NSLICES = NTOT / NCTA;        // 3 column slices of 256
constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // warpgroup 0: producer, MMA, 2 idle warps; warpgroups 1-2: epilogue
constexpr int CHUNK = 32 * 128;             // 32 rows x 128 bytes
constexpr int EPI_PER_WARP = 2 * CHUNK;     // pass A: two residual chunks; pass B: fp32 out chunk + bf16 out chunk
constexpr int STATS_BYTES = 2 * NSEGS * BM * 8;   // [slot][segment][row] float2
constexpr int VEC_BYTES = 3 * NCTA * 4;     // bias | gamma | beta of this CTA's 256 columns (current group)

// U2 = false: a cluster of 3 CTAs (one per 256-column slice) owns a 128-row block; every CTA runs its own M = 128 MMAs.
// U2 = true : a cluster of 6 CTAs = 3 column slices x a PAIR of CTAs (ranks 2n, 2n+1) owns a 256-row block; each pair runs
//             one tcgen05.mma.cta_group::2 (M = 256) per K step with half of the W slab per CTA, i.e. 32 KiB instead of
//             48 KiB taken in per K slab and a 4-stage ring -- the single-CTA form is bound by SM ingress (gemm.cu).
template <bool U2>
struct LnCfg {
    static constexpr int CLUSTER = U2 ? 6 : 3;
    static constexpr int ROWS = U2 ? 2 * BM : BM;                       // rows of a cluster's block
    static constexpr int STAGES = U2 ? 4 : 3;
    static constexpr int B_ROWS = U2 ? NCTA / 2 : NCTA;                 // W rows held per CTA and stage
    static constexpr int B_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int STATS_OFFSET = EPI_OFFSET + EPI_WARPS * EPI_PER_WARP;
    static constexpr int VEC_OFFSET = STATS_OFFSET + STATS_BYTES;
    static constexpr int BAR_OFFSET = VEC_OFFSET + VEC_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * EPI_WARPS + 2;   // full, empty, tfull[2], tempty[2], res[8][2], stats[2]
    static constexpr int SMEM_TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
    static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget exceeded");
};
constexpr int MAX_GROUPS = 4;

struct LnMaps {
    // r / c32: fp32 residual in / fp32 output (SPLIT: bf16 hi residual / bf16 lo output); rlo: bf16 lo residual (SPLIT only)
    CUtensorMap a[MAX_GROUPS], b[MAX_GROUPS], r[MAX_GROUPS], rlo[MAX_GROUPS], c32[MAX_GROUPS], c16[MAX_GROUPS];
};
struct LnParams {
    const float* bias[MAX_GROUPS];
    const float* gamma[MAX_GROUPS];
    const float* beta[MAX_GROUPS];
    int M[MAX_GROUPS];
    int blk_start[MAX_GROUPS + 1];   // first 128-row block of each group
    int groups;
    int K;
    float eps;
    int dbg;   // timing experiments only (MRA_LN_DEBUG, -DMRA_INSTRUMENT): 1 = no residual, 2 = no cross-CTA exchange,
               // 4 = no stores, 8 = clocks
};

__device__ unsigned long long g_ln_timing[16];   // MRA_LN_DEBUG & 8: cycles per epilogue phase (warp 2 / lane 0 of CTA 0)

__device__ __forceinline__ uint32_t swz(int r, int j) { return static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// 8-byte remote store into a peer CTA's shared memory; completes 8 bytes on that CTA's mbarrier when it has landed
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, uint64_t v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

__device__ __forceinline__ void decode_blk(const LnParams& p, int blk, int& g, int& m_blk) {
    g = 0;
#pragma unroll
    for (int i = 1; i < MAX_GROUPS; ++i)
        if (i < p.groups && blk >= p.blk_start[i]) g = i;
    m_blk = blk - p.blk_start[g];
}

template <bool U2, bool SPLIT>
__global__ void __cluster_dims__(LnCfg<U2>::CLUSTER, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ LnMaps maps, const __grid_constant__ LnParams p) {
    using C = LnCfg<U2>;
    constexpr int CLUSTER = C::CLUSTER, STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    // every CTA of the cluster computes the same offset (same kernel, same dynamic smem base): the distributed-shared-
    // memory addressing below relies on identical layouts
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float2* stats = reinterpret_cast<float2*>(smem + C::STATS_OFFSET);   // [2][NSEGS][BM]
    float* vecs = reinterpret_cast<float*>(smem + C::VEC_OFFSET);        // [bias|gamma|beta][256]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_bar = tempty_bar + 2;              // [EPI_WARPS][2]
    uint64_t* stats_bar = res_bar + 2 * EPI_WARPS;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stats_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t nidx = U2 ? rank >> 1 : rank;     // which 256-column slice
    const uint32_t mhalf = U2 ? rank & 1u : 0u;      // which 128-row half of the cluster's block (pair member)
    const uint32_t lead = U2 ? rank & ~1u : rank;    // rank of the pair's MMA issuer
    const int cluster_id = blockIdx.x / CLUSTER;
    const int num_clusters = gridDim.x / CLUSTER;
    const int total_blks = p.blk_start[p.groups];
    const int k_blocks = (p.K + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < p.groups; ++g) {
            ptx::prefetch_tensormap(&maps.a[g]);
            ptx::prefetch_tensormap(&maps.b[g]);
            ptx::prefetch_tensormap(&maps.r[g]);
            if (SPLIT) ptx::prefetch_tensormap(&maps.rlo[g]);
            ptx::prefetch_tensormap(&maps.c32[g]);
            ptx::prefetch_tensormap(&maps.c16[g]);
        }
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], U2 ? 2 * EPI_WARPS : EPI_WARPS);   // (U2: the epilogue warps of both pair members)
            ptx::mbar_init(&stats_bar[i], EPI_WARPS);   // + the peers' st.async bytes
        }
        for (int i = 0; i < 2 * EPI_WARPS; ++i) ptx::mbar_init(&res_bar[i], 1);
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
    cluster_sync_all();   // the peers' barriers are initialised before anyone signals them remotely
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above overlapped the tail of the previous kernel in the stream (programmatic dependent launch)
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // producer / MMA warpgroup gives up registers
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int blk = cluster_id; blk < total_blks; blk += num_clusters) {
                int g, m_blk;
                decode_blk(p, blk, g, m_blk);
                const CUtensorMap* tmA = &maps.a[g];
                const CUtensorMap* tmB = &maps.b[g];
                const int a_row = m_blk * C::ROWS + static_cast<int>(mhalf) * BM;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    if (U2) {
                        // both pair members' boxes complete on the leader's barrier, which expects the bytes of the pair
                        const uint32_t lead_bar = mapa(ptx::smem_u32(&full_bar[stage]), lead);
                        if (mhalf == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        ptx::tma_load_2d_pair(st, tmA, lead_bar, kb * BK, a_row);
                        ptx::tma_load_2d_pair(st + A_BYTES, tmB, lead_bar, kb * BK,
                                              static_cast<int>(nidx) * NCTA + static_cast<int>(mhalf) * C::B_ROWS);
                    } else {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        ptx::tma_load_2d(st, tmA, &full_bar[stage], kb * BK, a_row);
                        ptx::tma_load_2d(st + A_BYTES, tmB, &full_bar[stage], kb * BK, static_cast<int>(nidx) * NCTA);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0 && mhalf == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(U2 ? 2 * BM : BM, NCTA);
            const uint16_t pair_mask = static_cast<uint16_t>(0x3u << lead);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int blk = cluster_id; blk < total_blks; blk += num_clusters, ++iter) {
                const int acc = iter & 1;
                ptx::mbar_wait(&tempty_bar[acc], ((iter >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * NCTA;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = ptx::make_sw128_kmajor_desc(st);
                    const uint64_t b_desc = ptx::make_sw128_kmajor_desc(st + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (U2) ptx::umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                        else ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    if (U2) ptx::umma_commit_pair(&empty_bar[stage], pair_mask);
                    else ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (U2) ptx::umma_commit_pair(&tfull_bar[acc], pair_mask);
                else ptx::umma_commit(&tfull_bar[acc]);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue warps 4..11 (two warpgroups)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");   // 128 row values + staging per thread
        const int quad = warp & 3;            // TMEM lanes / tile rows [32 quad, 32 quad + 32)
        const int ew = warp - 4;
        const int member = ew >> 2;           // column segment of this CTA: [128 member, 128 member + 128)
        const int seg = static_cast<int>(nidx) * 2 + member;   // 0..5: position of the segment inside the 768 columns
        uint8_t* my = smem + C::EPI_OFFSET + ew * EPI_PER_WARP;
        const uint32_t my_s = ptx::smem_u32(my);
        uint64_t* rbar = res_bar + 2 * ew;
        uint32_t rphase = 0;
        const int row_in_tile = quad * 32 + lane;
        const uint32_t stats_local = ptx::smem_u32(stats);
        int cur_g = -1;
        int iter = 0;
        for (int blk = cluster_id; blk < total_blks; blk += num_clusters, ++iter) {
            int g, m_blk;
            decode_blk(p, blk, g, m_blk);
            const CUtensorMap* tmR = &maps.r[g];
            const CUtensorMap* tmRlo = &maps.rlo[g];
            const CUtensorMap* tmC32 = &maps.c32[g];
            const CUtensorMap* tmC16 = &maps.c16[g];
            const int Mg = p.M[g];
            const int acc = iter & 1;
            const int row0 = m_blk * C::ROWS + static_cast<int>(mhalf) * BM + quad * 32;
            const int col0 = seg * NSEG;
            const int slot = iter & 1;
            const bool no_res = MRA_LN_DBG(p) & 1, no_xchg = MRA_LN_DBG(p) & 2, no_store = MRA_LN_DBG(p) & 4;
            const bool prof = (MRA_LN_DBG(p) & 8) && blockIdx.x == 0 && ew == 0 && lane == 0;
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            if (prof) t0 = clock64();
            if (lane == 0 && !no_res) {
                // first residual chunk(s): issued before the accumulator is ready (overlaps the main loop)
                if constexpr (SPLIT) {
                    // 64 columns: hi box -> buffer 0, lo box -> buffer 1, both complete on rbar[0]
                    ptx::mbar_arrive_expect_tx(&rbar[0], 2 * CHUNK);
                    ptx::tma_load_2d(my, tmR, &rbar[0], col0, row0);
                    ptx::tma_load_2d(my + CHUNK, tmRlo, &rbar[0], col0, row0);
                } else {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        ptx::mbar_arrive_expect_tx(&rbar[b], CHUNK);
                        ptx::tma_load_2d(my + b * CHUNK, tmR, &rbar[b], col0 + b * 32, row0);
                    }
                }
            }
            if (g != cur_g) {
                // per-column vectors of this CTA's 256 columns -> shared memory (broadcast reads in the passes below)
                epi_bar_sync();   // nobody still reads the previous group's vectors
                for (int i = threadIdx.x - 128; i < 3 * NCTA; i += EPI_WARPS * 32) {
                    const int k = i / NCTA, cidx = i - k * NCTA;
                    const float* src = k == 0 ? p.bias[g] : (k == 1 ? p.gamma[g] : p.beta[g]);
                    vecs[i] = src != nullptr ? __ldg(src + nidx * NCTA + cidx) : 0.f;
                }
                epi_bar_sync();
                cur_g = g;
            }
            const float* bias = vecs + member * NSEG;
            const float* gamma = vecs + NCTA + member * NSEG;
            const float* beta = vecs + 2 * NCTA + member * NSEG;

            ptx::mbar_wait(&tfull_bar[acc], (iter >> 1) & 1);
            if (prof) t1 = clock64();
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * NCTA + member * NSEG;
            // ---- pass A: x = acc + bias + residual, kept in registers; row statistics
            float v[NSEG];
            float sum = 0.f, sumsq = 0.f;
            auto release_accumulator = [&]() {
                // last TMEM read of this block by this warp: the MMA warp may reuse the accumulator stage
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (U2) ptx::mbar_arrive_cluster(mapa(ptx::smem_u32(&tempty_bar[acc]), lead));   // the pair leader's barrier
                    else ptx::mbar_arrive(&tempty_bar[acc]);
                }
            };
            if constexpr (SPLIT) {
#pragma unroll
                for (int cc = 0; cc < NSEG / 64; ++cc) {
                    uint32_t r[2][32];
                    ptx::tmem_ld_32x32b_x32(t_row + cc * 64, r[0]);
                    ptx::tmem_ld_32x32b_x32(t_row + cc * 64 + 32, r[1]);
                    ptx::tmem_ld_wait();
                    if (cc == NSEG / 64 - 1) release_accumulator();
                    if (!no_res) {
                        ptx::mbar_wait(&rbar[0], rphase & 1u);
                        rphase ^= 1u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {   // 16-byte unit j of the row's 128-byte line: columns 8 j .. 8 j + 7 of the chunk
                        const uint4 hi = ptx::ld_shared_v4(my_s + swz(lane, j));
                        const uint4 lo = ptx::ld_shared_v4(my_s + CHUNK + swz(lane, j));
                        const float4 b0 = *reinterpret_cast<const float4*>(bias + cc * 64 + 8 * j);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias + cc * 64 + 8 * j + 4);
                        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int idx = 8 * j + 2 * e;   // column inside the 64-column chunk
                            const float a0 = __uint_as_float(r[idx >> 5][idx & 31]), a1 = __uint_as_float(r[(idx + 1) >> 5][(idx + 1) & 31]);
                            const float x0 = a0 + bb[2 * e] + (ptx::bf16lo(hw[e]) + ptx::bf16lo(lw[e]));
                            const float x1 = a1 + bb[2 * e + 1] + (ptx::bf16hi(hw[e]) + ptx::bf16hi(lw[e]));
                            sum += x0 + x1;
                            sumsq = fmaf(x0, x0, fmaf(x1, x1, sumsq));
                            v[cc * 64 + idx] = x0;
                            v[cc * 64 + idx + 1] = x1;
                        }
                    }
                    __syncwarp();   // every lane has read the residual buffers: refill them
                    if (lane == 0 && cc + 1 < NSEG / 64 && !no_res) {
                        ptx::mbar_arrive_expect_tx(&rbar[0], 2 * CHUNK);
                        ptx::tma_load_2d(my, tmR, &rbar[0], col0 + (cc + 1) * 64, row0);
                        ptx::tma_load_2d(my + CHUNK, tmRlo, &rbar[0], col0 + (cc + 1) * 64, row0);
                    }
                }
            } else {
                constexpr int NC = NSEG / 32;         // 4 chunks of 32 columns
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int b = c & 1;
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(t_row + c * 32, r);
                    ptx::tmem_ld_wait();
                    if (c == NC - 1) release_accumulator();
                    if (!no_res) {
                        ptx::mbar_wait(&rbar[b], (rphase >> b) & 1u);
                        rphase ^= 1u << b;
                    }
                    const uint32_t rs = my_s + b * CHUNK;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = ptx::ld_shared_v4f(rs + swz(lane, j));
                        const float4 bv = *reinterpret_cast<const float4*>(bias + c * 32 + 4 * j);
                        const float v0 = __uint_as_float(r[4 * j]) + bv.x + x.x;
                        const float v1 = __uint_as_float(r[4 * j + 1]) + bv.y + x.y;
                        const float v2 = __uint_as_float(r[4 * j + 2]) + bv.z + x.z;
                        const float v3 = __uint_as_float(r[4 * j + 3]) + bv.w + x.w;
                        sum += (v0 + v1) + (v2 + v3);
                        sumsq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, sumsq))));
                        v[c * 32 + 4 * j] = v0; v[c * 32 + 4 * j + 1] = v1; v[c * 32 + 4 * j + 2] = v2; v[c * 32 + 4 * j + 3] = v3;
                    }
                    __syncwarp();   // every lane has read the residual buffer: refill it
                    if (lane == 0 && c + 2 < NC && !no_res) {
                        ptx::mbar_arrive_expect_tx(&rbar[b], CHUNK);
                        ptx::tma_load_2d(my + b * CHUNK, tmR, &rbar[b], col0 + (c + 2) * 32, row0);
                    }
                }
            }
            if (prof) t2 = clock64();
            // ---- exchange the partial statistics: own stats buffer + the two peers'
            {
                const uint32_t off = static_cast<uint32_t>(((slot * NSEGS + seg) * BM + row_in_tile) * 8);
                stats[(slot * NSEGS + seg) * BM + row_in_tile] = make_float2(sum, sumsq);
                const uint64_t packed = (static_cast<uint64_t>(__float_as_uint(sumsq)) << 32) | __float_as_uint(sum);
                if (!no_xchg) {
#pragma unroll
                    for (uint32_t d = 1; d < NSLICES; ++d) {
                        // the CTAs holding the other column slices of the SAME rows
                        const uint32_t pn = (nidx + d) % NSLICES;
                        const uint32_t peer = U2 ? 2 * pn + mhalf : pn;
                        st_async_b64(mapa(stats_local + off, peer), packed, mapa(ptx::smem_u32(&stats_bar[slot]), peer));
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    if (ew == 0 && !no_xchg) ptx::mbar_arrive_expect_tx(&stats_bar[slot], (NSLICES - 1) * EPI_WARPS * 32 * 8);
                    else ptx::mbar_arrive(&stats_bar[slot]);
                }
                ptx::mbar_wait(&stats_bar[slot], (iter >> 1) & 1);
            }
            float mean, rstd;
            {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int q = 0; q < NSEGS; ++q) {
                    const float2 sv = stats[(slot * NSEGS + q) * BM + row_in_tile];
                    s += sv.x;
                    ss += sv.y;
                }
                mean = s * (1.0f / NTOT);
                const float var = fmaxf(ss * (1.0f / NTOT) - mean * mean, 0.f);
                rstd = rsqrtf(var + p.eps);
            }
            if (prof) t3 = clock64();
            // ---- pass B: normalise from registers, scale / shift, write the output copies
            const float nm = -mean * rstd;
            if constexpr (SPLIT) {
                // 64-column chunks: buffer 0 = hi (bf16(y)), buffer 1 = lo (bf16(y - hi)); one 128-byte line per row each
#pragma unroll
                for (int cc = 0; cc < NSEG / 64; ++cc) {
                    if (cc > 0) {   // (chunk 0: the buffers held residual chunks whose loads were waited for above)
                        if (lane == 0) ptx::tma_store_wait_read<0>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 g0 = *reinterpret_cast<const float4*>(gamma + cc * 64 + 8 * j);
                        const float4 g1 = *reinterpret_cast<const float4*>(gamma + cc * 64 + 8 * j + 4);
                        const float4 e0 = *reinterpret_cast<const float4*>(beta + cc * 64 + 8 * j);
                        const float4 e1 = *reinterpret_cast<const float4*>(beta + cc * 64 + 8 * j + 4);
                        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                        const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float y0 = fmaf(fmaf(v[cc * 64 + 8 * j + 2 * e], rstd, nm), gg[2 * e], ee[2 * e]);
                            const float y1 = fmaf(fmaf(v[cc * 64 + 8 * j + 2 * e + 1], rstd, nm), gg[2 * e + 1], ee[2 * e + 1]);
                            hw[e] = ptx::pack_bf16x2(y0, y1);
                            lw[e] = ptx::pack_bf16x2(y0 - ptx::bf16lo(hw[e]), y1 - ptx::bf16hi(hw[e]));
                        }
                        ptx::st_shared_v4(my_s + swz(lane, j), hw[0], hw[1], hw[2], hw[3]);
                        ptx::st_shared_v4(my_s + CHUNK + swz(lane, j), lw[0], lw[1], lw[2], lw[3]);
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < Mg && !no_store) {
                            ptx::tma_store_2d(tmC16, my, col0 + cc * 64, row0);
                            ptx::tma_store_2d(tmC32, my + CHUNK, col0 + cc * 64, row0);
                        }
                        ptx::tma_store_commit();
                    }
                }
            } else {
                //      staging: buffer 1 = fp32 chunk (32 columns), buffer 0 = bf16 chunk (64 columns, stored every 2nd chunk)
                constexpr int NC = NSEG / 32;
                uint8_t* o16 = my;
                uint8_t* o32 = my + CHUNK;
                const uint32_t o16_s = my_s, o32_s = my_s + CHUNK;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (lane == 0) ptx::tma_store_wait_read<0>();   // the stores that last read the staging buffers are done
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 gv = *reinterpret_cast<const float4*>(gamma + c * 32 + 4 * j);
                        const float4 bv = *reinterpret_cast<const float4*>(beta + c * 32 + 4 * j);
                        const float y0 = fmaf(fmaf(v[c * 32 + 4 * j], rstd, nm), gv.x, bv.x);
                        const float y1 = fmaf(fmaf(v[c * 32 + 4 * j + 1], rstd, nm), gv.y, bv.y);
                        const float y2 = fmaf(fmaf(v[c * 32 + 4 * j + 2], rstd, nm), gv.z, bv.z);
                        const float y3 = fmaf(fmaf(v[c * 32 + 4 * j + 3], rstd, nm), gv.w, bv.w);
                        ptx::st_shared_v4f(o32_s + swz(lane, j), y0, y1, y2, y3);
                        // 8 bytes of bf16 at column 4 j of this chunk: 16-byte unit (c & 1) * 4 + j / 2, half j & 1
                        ptx::st_shared_v2(o16_s + swz(lane, (c & 1) * 4 + (j >> 1)) + (j & 1) * 8, ptx::pack_bf16x2(y0, y1),
                                          ptx::pack_bf16x2(y2, y3));
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (row0 < Mg && !no_store) {
                            ptx::tma_store_2d(tmC32, o32, col0 + c * 32, row0);
                            if (c & 1) ptx::tma_store_2d(tmC16, o16, col0 + (c - 1) * 32, row0);
                        }
                        ptx::tma_store_commit();
                    }
                }
            }
            if (prof) t4 = clock64();
            // the staging buffers take the next block's residual chunks: every store must have finished reading them
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
            if (prof) {
                const long long t5 = clock64();
                g_ln_timing[0] += t1 - t0; g_ln_timing[1] += t2 - t1; g_ln_timing[2] += t3 - t2;
                g_ln_timing[3] += t4 - t3; g_ln_timing[4] += t5 - t4; g_ln_timing[5] += 1;
            }
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA exits while a peer may still write into its shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        if (U2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <bool U2, bool SPLIT>
static int launch_gemm_ln_variant(const GemmLnArgs* ga, int n, float eps, cudaStream_t s) {
    using C = LnCfg<U2>;
    auto kern = gemm_ln_kernel<U2, SPLIT>;
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(kern), C::SMEM_TOTAL)) return e;
    LnMaps maps;
    LnParams p;
    p.groups = n;
    p.K = ga[0].K;
    p.eps = eps;
    static const int dbg = [] { const char* e = getenv("MRA_LN_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
    int total = 0;
    for (int g = 0; g < MAX_GROUPS; ++g) {
        if (g < n) {
            const GemmLnArgs& a = ga[g];
            MRA_REQUIRE(a.M > 0 && a.K > 0 && a.K % 8 == 0 && a.K == ga[0].K, "fused GEMM+LayerNorm: bad / mismatching K");
            MRA_REQUIRE(a.A && a.W && a.residual && a.gamma && a.beta && a.y16, "fused GEMM+LayerNorm: NULL operand");
            MRA_REQUIRE(SPLIT ? (a.res_lo && a.y_lo) : (a.y32 != nullptr), "fused GEMM+LayerNorm: NULL residual / output operand");
            if (int e = get_tensor_map(a.A, a.M, a.K, a.lda, BM, BK, 2, &maps.a[g])) return e;
            if (int e = get_tensor_map(a.W, NTOT, a.K, a.ldw, C::B_ROWS, BK, 2, &maps.b[g])) return e;
            if (SPLIT) {
                if (int e = get_tensor_map(a.residual, a.M, NTOT, a.ldr, 32, 64, 2, &maps.r[g])) return e;
                if (int e = get_tensor_map(a.res_lo, a.M, NTOT, a.ldr, 32, 64, 2, &maps.rlo[g])) return e;
                if (int e = get_tensor_map(a.y_lo, a.M, NTOT, a.ldy16, 32, 64, 2, &maps.c32[g])) return e;
            } else {
                if (int e = get_tensor_map(a.residual, a.M, NTOT, a.ldr, 32, 32, 4, &maps.r[g])) return e;
                maps.rlo[g] = maps.r[g];
                if (int e = get_tensor_map(a.y32, a.M, NTOT, a.ldy32, 32, 32, 4, &maps.c32[g])) return e;
            }
            if (int e = get_tensor_map(a.y16, a.M, NTOT, a.ldy16, 32, 64, 2, &maps.c16[g])) return e;
            p.bias[g] = a.bias; p.gamma[g] = a.gamma; p.beta[g] = a.beta;
            p.M[g] = a.M;
            p.blk_start[g] = total;
            total += (a.M + C::ROWS - 1) / C::ROWS;
        } else {
            maps.a[g] = maps.a[0]; maps.b[g] = maps.b[0]; maps.r[g] = maps.r[0]; maps.rlo[g] = maps.rlo[0];
            maps.c32[g] = maps.c32[0]; maps.c16[g] = maps.c16[0];
            p.bias[g] = p.gamma[g] = p.beta[g] = nullptr;
            p.M[g] = 0;
            p.blk_start[g] = total;
        }
    }
    for (int g = n; g <= MAX_GROUPS; ++g) p.blk_start[g] = total;
    // persistent grid = as many clusters as can be co-resident (3- / 6-CTA clusters do not tile every GPC completely)
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C::CLUSTER * (sm_count() / C::CLUSTER));
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_TOTAL;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C::CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc <= 0) nc = sm_count() / C::CLUSTER;
        max_clusters = nc;
        if (getenv("MRA_LN_DEBUG")) fprintf(stderr, "[gemm_ln] max co-resident %d-CTA clusters: %d\n", C::CLUSTER, nc);
    }
    int clusters = max_clusters;
    if (clusters > total) clusters = total;
    // (the cluster size is the kernel's compile-time __cluster_dims__: no launch attribute for it)
    MRA_CHECK_CUDA(launch_pdl(kern, dim3(C::CLUSTER * clusters), dim3(NUM_THREADS), C::SMEM_TOTAL, s, 1, maps, p));
#ifdef MRA_INSTRUMENT
    if (dbg & 8) {
        unsigned long long t[16];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(t, g_ln_timing, sizeof(t));
        const double nt = t[5] ? double(t[5]) : 1.0;
        fprintf(stderr, "[gemm_ln U2=%d SPLIT=%d timing, cycles/tile over %llu tiles] wait-mainloop %.0f | pass1 %.0f | exchange %.0f | pass2 %.0f | drain %.0f\n",
                int(U2), int(SPLIT), t[5], t[0] / nt, t[1] / nt, t[2] / nt, t[3] / nt, t[4] / nt);
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_ln_timing, z, sizeof(z));
    }
#endif
    return 0;
}

}  // namespace

int launch_gemm_ln_grouped(const GemmLnArgs* ga, int n, float eps, cudaStream_t s) {
    MRA_REQUIRE(n >= 1 && n <= MAX_GROUPS, "fused GEMM+LayerNorm takes 1..%d problems, got %d", MAX_GROUPS, n);
    const bool split = ga[0].res_lo != nullptr;
    for (int g = 1; g < n; ++g)
        MRA_REQUIRE((ga[g].res_lo != nullptr) == split, "fused GEMM+LayerNorm: grouped problems must share the residual form");
    // 1 (default) = 6-CTA clusters with 2-CTA MMAs, 0 = 3-CTA clusters with single-CTA MMAs (A/B runs, tests)
    static const bool u2 = [] { const char* e = getenv("MRA_LN_U2"); return e == nullptr || atoi(e) != 0; }();
    if (split) return u2 ? launch_gemm_ln_variant<true, true>(ga, n, eps, s) : launch_gemm_ln_variant<false, true>(ga, n, eps, s);
    return u2 ? launch_gemm_ln_variant<true, false>(ga, n, eps, s) : launch_gemm_ln_variant<false, false>(ga, n, eps, s);
}

}  // namespace mra
