// Y = LayerNorm(A . W^T + bias + residual) * gamma + beta   for N = 768 (the Q-Former hidden size) in ONE kernel:
// the out-projection / FFN-down Linear of every Q-Former block together with its residual add and post-LayerNorm
// (BertSelfOutput / BertOutput, HF port modeling_instructblip.py:549-553, 606-610).  The pre-LayerNorm sums never
// touch HBM: they stay in TMEM between the GEMM main loop and the normalisation.
//
// A 128 x 768 fp32 accumulator needs 768 TMEM columns (an SM has 512), so a thread-block CLUSTER of two CTAs owns one
// 128-row block: CTA r computes columns [384 r, 384 r + 384) with two tcgen05.mma (M=128, N=192) per 16-wide K step,
// and the per-row LayerNorm statistics are combined across the pair through distributed shared memory
// (st.async ... mbarrier::complete_tx into the peer's stats buffer).  Per CTA:
//   warp 0      TMA producer: A tile 128 x 64 + W tile 384 x 64 per stage (64 KiB), 3 stages
//   warp 1      MMA issuer (single thread), one TMEM accumulator of 384 columns
//   warps 4..11 epilogue (two warpgroups, 232 registers per thread via setmaxnreg), thread = (row, 192-column segment):
//       pass A  tcgen05.ld (once) -> + bias + residual (fp32, TMA-loaded chunks), values stay in registers; row sum / sumsq
//       exchange partial statistics inside the CTA and with the peer CTA
//       pass B  normalise from registers, * gamma + beta -> fp32 and bf16 copies through swizzled staging + TMA stores
// The epilogue's staging buffers live in pipeline stages 1 and 2, which are idle while the accumulator is being
// drained (there is a single accumulator, so the next tile's main loop cannot start anyway); the producer prefetches
// only the next tile's first K slab (stage 0) meanwhile.  Up to 4 grouped problems per launch (see gemm.cu).
#include "common.h"
#include "ptx.cuh"

namespace mra {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NTOT = 768;                   // LayerNorm width handled by a CTA pair
constexpr int NCTA = 384;                   // columns per CTA
constexpr int NSEG = 192;                   // columns per epilogue thread / per MMA
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
constexpr int B_HALF_BYTES = NSEG * BK * 2; // 24 KiB
constexpr int STAGE_BYTES = A_BYTES + 2 * B_HALF_BYTES;   // 64 KiB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;   // warpgroup 0: producer, MMA, 2 idle warps; warpgroups 1-2: epilogue
constexpr int CHUNK = 32 * 128;             // 32 rows x 128 bytes
constexpr int EPI_PER_WARP = 4 * CHUNK;     // R0, R1 (fp32 in / out), H0, H1 (bf16 out)  = 16 KiB
constexpr int STATS_BYTES = 2 * 4 * BM * 8; // [slot][source segment 0..3][row] float2
constexpr int VEC_BYTES = 4 * 3 * NCTA * 4;      // bias / gamma / beta of this CTA's 384 columns for up to 4 groups
constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + STATS_BYTES + VEC_BYTES;
constexpr int NUM_BARS = 2 * STAGES + 2 + 1 + 4 * EPI_WARPS + 2;   // full, empty, tfull, tempty, epi_done, res[8][4], stats[2]
constexpr int SMEM_TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
static_assert(2 * STAGE_BYTES >= EPI_WARPS * EPI_PER_WARP, "epilogue staging must fit into stages 1-2");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget exceeded");
constexpr int MAX_GROUPS = 4;

struct LnMaps {
    CUtensorMap a[MAX_GROUPS], b[MAX_GROUPS], r[MAX_GROUPS], c32[MAX_GROUPS], c16[MAX_GROUPS];
    CUtensorMap rp[MAX_GROUPS];   // residual again, box {32 cols, 128 rows}: L2 prefetch of a whole tile column block
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
    int dbg;   // timing experiments only (MRA_LN_DEBUG): 1 = no residual, 2 = no cross-CTA exchange, 4 = no stores
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
// 8-byte remote store into the peer CTA's shared memory; completes `bytes` on the peer's mbarrier when it has landed
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, uint64_t v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v),
                 "r"(remote_bar)
                 : "memory");
}
// bring a residual box into L2 ahead of the epilogue's TMA loads (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const void* desc, int32_t crd0, int32_t crd1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(crd0),
                 "r"(crd1)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void decode_blk(const LnParams& p, int blk, int& g, int& m_blk) {
    g = 0;
#pragma unroll
    for (int i = 1; i < MAX_GROUPS; ++i)
        if (i < p.groups && blk >= p.blk_start[i]) g = i;
    m_blk = blk - p.blk_start[g];
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_ln_kernel(const __grid_constant__ LnMaps maps, const __grid_constant__ LnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // NOTE: both CTAs of the cluster compute the same offset (same kernel, same dynamic smem base), which the
    // distributed-shared-memory addressing below relies on.
    float2* stats = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES);   // [2][4][BM]
    float* vecs = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + STATS_BYTES);   // [group][bias|gamma|beta][384]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 1;
    uint64_t* epi_done_bar = tempty_bar + 1;
    uint64_t* res_bar = epi_done_bar + 1;        // [EPI_WARPS][4]
    uint64_t* stats_bar = res_bar + 4 * EPI_WARPS;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stats_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int total_blks = p.blk_start[p.groups];
    const int k_blocks = (p.K + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = 512;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < p.groups; ++g) {
            ptx::prefetch_tensormap(&maps.a[g]);
            ptx::prefetch_tensormap(&maps.b[g]);
            ptx::prefetch_tensormap(&maps.r[g]);
            ptx::prefetch_tensormap(&maps.c32[g]);
            ptx::prefetch_tensormap(&maps.c16[g]);
        }
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(tempty_bar, EPI_WARPS);
        ptx::mbar_init(epi_done_bar, EPI_WARPS);
        for (int i = 0; i < 4 * EPI_WARPS; ++i) ptx::mbar_init(&res_bar[i], 1);
        for (int i = 0; i < 2; ++i) ptx::mbar_init(&stats_bar[i], EPI_WARPS);   // + the peer's 2 KiB of st.async bytes
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    // per-column vectors of this CTA's 384 columns -> shared memory (the epilogue reads them as broadcasts instead of
    // paying an L2 round trip per 32-column chunk)
    for (int i = threadIdx.x; i < p.groups * 3 * NCTA; i += NUM_THREADS) {
        const int g = i / (3 * NCTA), k = (i / NCTA) % 3, cidx = i % NCTA;
        const float* src = k == 0 ? p.bias[g] : (k == 1 ? p.gamma[g] : p.beta[g]);
        vecs[i] = src != nullptr ? __ldg(src + cluster_ctarank() * NCTA + cidx) : 0.f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    cluster_sync_all();   // the peer's barriers are initialised before anyone signals them remotely
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            uint32_t ephase = 0;   // bit s = parity of the next wait on empty_bar[s]  (starts "free")
            int iter = 0;
            for (int blk = cluster_id; blk < total_blks; blk += num_clusters, ++iter) {
                int g, m_blk;
                decode_blk(p, blk, g, m_blk);
                const CUtensorMap* tmA = &maps.a[g];
                const CUtensorMap* tmB = &maps.b[g];
                if (!(p.dbg & 16)) {
                    // the epilogue will read this tile's residual block right after the main loop: start moving it into L2
#pragma unroll 1
                    for (int c = 0; c < NCTA / 32; ++c)
                        tma_prefetch_2d(&maps.rp[g], static_cast<int>(rank) * NCTA + c * 32, m_blk * BM);
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int stage = kb % STAGES;
                    if (kb == 1 && iter > 0) {
                        // stages 1-2 double as the epilogue's staging area: wait until the previous tile is drained
                        ptx::mbar_wait(epi_done_bar, (iter - 1) & 1);
                    }
                    ptx::mbar_wait(&empty_bar[stage], ((ephase >> stage) & 1u) ^ 1u);
                    ephase ^= 1u << stage;
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                    ptx::tma_load_2d(st, tmA, &full_bar[stage], kb * BK, m_blk * BM);
                    ptx::tma_load_2d(st + A_BYTES, tmB, &full_bar[stage], kb * BK, static_cast<int>(rank) * NCTA);
                    ptx::tma_load_2d(st + A_BYTES + B_HALF_BYTES, tmB, &full_bar[stage], kb * BK, static_cast<int>(rank) * NCTA + NSEG);
                }
                if (k_blocks == 1 && iter > 0) ptx::mbar_wait(epi_done_bar, (iter - 1) & 1);   // keep the phases in step
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, NSEG);
            uint32_t fphase = 0;   // bit s = parity of the next wait on full_bar[s]
            int iter = 0;
            for (int blk = cluster_id; blk < total_blks; blk += num_clusters, ++iter) {
                ptx::mbar_wait(tempty_bar, (iter & 1) ^ 1);   // the epilogue has drained the accumulator
                ptx::tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int stage = kb % STAGES;
                    ptx::mbar_wait(&full_bar[stage], (fphase >> stage) & 1u);
                    fphase ^= 1u << stage;
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = ptx::make_sw128_kmajor_desc(st);
                    const uint64_t b0_desc = ptx::make_sw128_kmajor_desc(st + A_BYTES);
                    const uint64_t b1_desc = ptx::make_sw128_kmajor_desc(st + A_BYTES + B_HALF_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        ptx::umma_bf16_ss(tmem_base, a_desc + 2 * k, b0_desc + 2 * k, idesc, (kb | k) != 0);
                        ptx::umma_bf16_ss(tmem_base + NSEG, a_desc + 2 * k, b1_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                }
                ptx::umma_commit(tfull_bar);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue warps 4..11 (two warpgroups)
        // TMEM reads cost 64 B/clk per SM (3072 cycles for this CTA's 128 x 384 fp32 block), so the block is read ONCE:
        // every thread keeps its 192 pre-LayerNorm values in registers (the warpgroups take the registers the
        // producer / MMA warpgroup gives up) and hands the accumulator back to the MMA warp before normalising.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int quad = warp & 3;            // TMEM lanes / tile rows [32 quad, 32 quad + 32)
        const int ew = warp - 4;
        const int member = ew >> 2;           // column segment of this CTA: [192 member, 192 member + 192)
        const int seg = static_cast<int>(rank) * 2 + member;   // 0..3: position of the segment inside the 768 columns
        uint8_t* my = smem + STAGE_BYTES + ew * EPI_PER_WARP;   // inside stages 1-2
        uint64_t* rbar = res_bar + 4 * ew;
        uint32_t rphase = 0;
        const int row_in_tile = quad * 32 + lane;
        const uint32_t stats_local = ptx::smem_u32(stats);
        const uint32_t peer = rank ^ 1u;
        constexpr int NC = NSEG / 32;         // 6 chunks of 32 columns
        int iter = 0;
        for (int blk = cluster_id; blk < total_blks; blk += num_clusters, ++iter) {
            int g, m_blk;
            decode_blk(p, blk, g, m_blk);
            const CUtensorMap* tmR = &maps.r[g];
            const CUtensorMap* tmC32 = &maps.c32[g];
            const CUtensorMap* tmC16 = &maps.c16[g];
            const float* bias = vecs + (g * 3 + 0) * NCTA + member * NSEG;    // this thread's 192-column segment
            const float* gamma = vecs + (g * 3 + 1) * NCTA + member * NSEG;
            const float* beta = vecs + (g * 3 + 2) * NCTA + member * NSEG;
            const int Mg = p.M[g];
            const int row0 = m_blk * BM + quad * 32;
            const int col0 = seg * NSEG;
            const int slot = iter & 1;
            const bool no_res = p.dbg & 1, no_xchg = p.dbg & 2, no_store = p.dbg & 4;
            const bool prof = (p.dbg & 8) && blockIdx.x == 0 && ew == 0 && lane == 0;
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
            if (prof) t0 = clock64();
            ptx::mbar_wait(tfull_bar, iter & 1);   // accumulator complete => stages 1-2 are no longer read by the MMAs
            if (prof) t1 = clock64();
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + member * NSEG;
            if (lane == 0 && !no_res) {
                // all four staging buffers take residual chunks during pass A (the bf16 buffers are idle until pass B)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    ptx::mbar_arrive_expect_tx(&rbar[b], CHUNK);
                    ptx::tma_load_2d(my + b * CHUNK, tmR, &rbar[b], col0 + b * 32, row0);
                }
            }
            // ---- pass A: x = acc + bias + residual, kept in registers; row statistics
            float v[NC][32];
            float sum = 0.f, sumsq = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int b = c & 3;
                uint32_t r[32];
                long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                if (prof) a0 = clock64();
                ptx::tmem_ld_32x32b_x32(t_row + c * 32, r);
                ptx::tmem_ld_wait();
                if (prof) a1 = clock64();
                if (c == NC - 1) {
                    // last TMEM read of this tile by this warp: the MMA warp may start the next tile's main loop
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty_bar);
                }
                if (!no_res) {
                    ptx::mbar_wait(&rbar[b], (rphase >> b) & 1u);
                    rphase ^= 1u << b;
                }
                if (prof) a2 = clock64();
                const uint8_t* rs = my + b * CHUNK;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(rs + swz(lane, j));
                    const float4 bv = *reinterpret_cast<const float4*>(bias + c * 32 + 4 * j);
                    const float v0 = __uint_as_float(r[4 * j]) + bv.x + x.x;
                    const float v1 = __uint_as_float(r[4 * j + 1]) + bv.y + x.y;
                    const float v2 = __uint_as_float(r[4 * j + 2]) + bv.z + x.z;
                    const float v3 = __uint_as_float(r[4 * j + 3]) + bv.w + x.w;
                    sum += (v0 + v1) + (v2 + v3);
                    sumsq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, sumsq))));
                    v[c][4 * j] = v0; v[c][4 * j + 1] = v1; v[c][4 * j + 2] = v2; v[c][4 * j + 3] = v3;
                }
                __syncwarp();   // every lane has read the residual buffer: refill it
                if (lane == 0 && c + 4 < NC && !no_res) {
                    ptx::mbar_arrive_expect_tx(&rbar[b], CHUNK);
                    ptx::tma_load_2d(my + b * CHUNK, tmR, &rbar[b], col0 + (c + 4) * 32, row0);
                }
                if (prof) {
                    a3 = clock64();
                    g_ln_timing[6] += a1 - a0; g_ln_timing[7] += a2 - a1; g_ln_timing[8] += a3 - a2;
                }
            }
            if (prof) t2 = clock64();
            // ---- exchange the partial statistics: local segment -> own stats buffer and the peer's
            {
                const uint32_t off = static_cast<uint32_t>(((slot * 4 + seg) * BM + row_in_tile) * 8);
                stats[(slot * 4 + seg) * BM + row_in_tile] = make_float2(sum, sumsq);
                const uint64_t packed = (static_cast<uint64_t>(__float_as_uint(sumsq)) << 32) | __float_as_uint(sum);
                if (!no_xchg) st_async_b64(mapa(stats_local + off, peer), packed, mapa(ptx::smem_u32(&stats_bar[slot]), peer));
                __syncwarp();
                if (lane == 0) {
                    if (ew == 0 && !no_xchg) ptx::mbar_arrive_expect_tx(&stats_bar[slot], EPI_WARPS * 32 * 8);   // the peer's 256 x 8 bytes
                    else ptx::mbar_arrive(&stats_bar[slot]);
                }
                ptx::mbar_wait(&stats_bar[slot], (iter >> 1) & 1);
            }
            float mean, rstd;
            {
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 sv = stats[(slot * 4 + q) * BM + row_in_tile];
                    s += sv.x;
                    ss += sv.y;
                }
                mean = s * (1.0f / NTOT);
                const float var = fmaxf(ss * (1.0f / NTOT) - mean * mean, 0.f);
                rstd = rsqrtf(var + p.eps);
            }
            if (prof) t3 = clock64();
            // ---- pass B: normalise from registers, scale / shift, write fp32 + bf16 copies
            const float nm = -mean * rstd;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int b = c & 1;
                uint8_t* o32 = my + b * CHUNK;
                uint8_t* o16 = my + (2 + ((c >> 1) & 1)) * CHUNK;   // one bf16 buffer per pair of chunks (64 columns)
                long long b0 = 0, b1 = 0, b2 = 0, b3 = 0;
                if (prof) b0 = clock64();
                if (lane == 0) ptx::tma_store_wait_read<1>();      // the stores that last read these buffers are done reading
                __syncwarp();
                if (prof) b1 = clock64();
                uint32_t h16[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 gv = *reinterpret_cast<const float4*>(gamma + c * 32 + 4 * j);
                    const float4 bv = *reinterpret_cast<const float4*>(beta + c * 32 + 4 * j);
                    const float y0 = fmaf(fmaf(v[c][4 * j], rstd, nm), gv.x, bv.x);
                    const float y1 = fmaf(fmaf(v[c][4 * j + 1], rstd, nm), gv.y, bv.y);
                    const float y2 = fmaf(fmaf(v[c][4 * j + 2], rstd, nm), gv.z, bv.z);
                    const float y3 = fmaf(fmaf(v[c][4 * j + 3], rstd, nm), gv.w, bv.w);
                    *reinterpret_cast<float4*>(o32 + swz(lane, j)) = make_float4(y0, y1, y2, y3);
                    h16[2 * j] = ptx::pack_bf16x2(y0, y1);
                    h16[2 * j + 1] = ptx::pack_bf16x2(y2, y3);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(o16 + swz(lane, (c & 1) * 4 + j)) =
                        make_uint4(h16[4 * j], h16[4 * j + 1], h16[4 * j + 2], h16[4 * j + 3]);
                if (prof) b2 = clock64();
                ptx::fence_proxy_async();
                __syncwarp();
                if (prof) b3 = clock64();
                if (lane == 0) {
                    if (row0 < Mg && !no_store) {
                        ptx::tma_store_2d(tmC32, o32, col0 + c * 32, row0);
                        if (c & 1) ptx::tma_store_2d(tmC16, o16, col0 + (c - 1) * 32, row0);
                    }
                    ptx::tma_store_commit();
                }
                if (prof) {
                    const long long b4 = clock64();
                    g_ln_timing[9] += b1 - b0; g_ln_timing[10] += b2 - b1; g_ln_timing[11] += b3 - b2; g_ln_timing[12] += b4 - b3;
                }
            }
            if (prof) t4 = clock64();
            // ---- the staging buffers go back to the producer once every store has finished reading them
            if (lane == 0) {
                ptx::tma_store_wait_read<0>();
                ptx::mbar_arrive(epi_done_bar);
            }
            if (prof) {
                const long long t5 = clock64();
                g_ln_timing[0] += t1 - t0; g_ln_timing[1] += t2 - t1; g_ln_timing[2] += t3 - t2;
                g_ln_timing[3] += t4 - t3; g_ln_timing[4] += t5 - t4; g_ln_timing[5] += 1;
            }
            __syncwarp();
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // warps 2, 3: idle members of the producer / MMA warpgroup
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA exits while its peer may still write into its shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

int launch_gemm_ln_grouped(const GemmLnArgs* ga, int n, float eps, cudaStream_t s) {
    MRA_REQUIRE(n >= 1 && n <= MAX_GROUPS, "fused GEMM+LayerNorm takes 1..%d problems, got %d", MAX_GROUPS, n);
    static bool attr_set = false;
    if (!attr_set) {
        MRA_CHECK_CUDA(cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        attr_set = true;
    }
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
            MRA_REQUIRE(a.A && a.W && a.residual && a.gamma && a.beta && a.y32 && a.y16, "fused GEMM+LayerNorm: NULL operand");
            if (int e = get_tensor_map(a.A, a.M, a.K, a.lda, BM, BK, 2, &maps.a[g])) return e;
            if (int e = get_tensor_map(a.W, NTOT, a.K, a.ldw, NSEG, BK, 2, &maps.b[g])) return e;
            if (int e = get_tensor_map(a.residual, a.M, NTOT, a.ldr, 32, 32, 4, &maps.r[g])) return e;
            if (int e = get_tensor_map(a.residual, a.M, NTOT, a.ldr, BM, 32, 4, &maps.rp[g])) return e;
            if (int e = get_tensor_map(a.y32, a.M, NTOT, a.ldy32, 32, 32, 4, &maps.c32[g])) return e;
            if (int e = get_tensor_map(a.y16, a.M, NTOT, a.ldy16, 32, 64, 2, &maps.c16[g])) return e;
            p.bias[g] = a.bias; p.gamma[g] = a.gamma; p.beta[g] = a.beta;
            p.M[g] = a.M;
            p.blk_start[g] = total;
            total += (a.M + BM - 1) / BM;
        } else {
            maps.a[g] = maps.a[0]; maps.b[g] = maps.b[0]; maps.r[g] = maps.r[0]; maps.c32[g] = maps.c32[0]; maps.c16[g] = maps.c16[0]; maps.rp[g] = maps.rp[0];
            p.bias[g] = p.gamma[g] = p.beta[g] = nullptr;
            p.M[g] = 0;
            p.blk_start[g] = total;
        }
    }
    for (int g = n; g <= MAX_GROUPS; ++g) p.blk_start[g] = total;
    int clusters = sm_count() / 2;
    if (clusters > total) clusters = total;
    gemm_ln_kernel<<<2 * clusters, NUM_THREADS, SMEM_TOTAL, s>>>(maps, p);
    MRA_CHECK_CUDA(cudaGetLastError());
    if (dbg & 8) {
        unsigned long long t[16];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(t, g_ln_timing, sizeof(t));
        const double n = t[5] ? double(t[5]) : 1.0;
        fprintf(stderr, "[gemm_ln timing, cycles/tile over %llu tiles] wait-mainloop %.0f | pass1 %.0f | exchange %.0f | pass2 %.0f | drain %.0f\n",
                t[5], t[0] / n, t[1] / n, t[2] / n, t[3] / n, t[4] / n);
        fprintf(stderr, "   passA per tile: tmem-ld %.0f | residual wait %.0f | math+refill %.0f ;  passB per tile: wait_read %.0f | math+sts %.0f | fence %.0f | issue %.0f\n",
                t[6] / n, t[7] / n, t[8] / n, t[9] / n, t[10] / n, t[11] / n, t[12] / n);
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_ln_timing, z, sizeof(z));
    }
    return 0;
}

}  // namespace mra
