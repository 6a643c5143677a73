// ctx = SelfAttention(x W_qkv^T + b_qkv)  -- the fused Q/K/V projection of a Q-Former block TOGETHER with its attention core
// (BertSelfAttention, HF port modeling_instructblip.py:499-538) in ONE kernel, for the geometry of the X-InstructBLIP step:
// 32 query tokens + 32 text tokens per row ("clip"), heads of 64, hidden = heads * 64.
//
// Why: unfused, the projection writes qkv [tokens, 3 H] (151 MB per layer at config 2) and the attention kernel reads it
// back; both launches then run at the rate at which HBM takes those bytes (profiles/r02_NOTES.md, session 4: the attention
// launches do not react to occupancy, ring depth, store width or operand layout).  Here Q, K and V of a tile never leave the
// SM: accumulator (TMEM) -> registers -> + bias -> bf16 -> shared memory -> attention on the epilogue warps -> ctx.
//
// Work item = (problem, block of 4 clips, head).  A pair of CTAs (cluster of 2) runs ONE tcgen05.mma.cta_group::2 per K
// step with M = 256 (4 clips x 64 tokens; each CTA holds the 2 x 64 token rows of two clips) and N = 192 (Q_h | K_h | V_h):
//   warp 0      TMA producer: per 64-wide K slab and CTA four 32-row boxes of x (clip a queries, clip a text, clip b queries,
//               clip b text -- the "split" token layout keeps all query rows before all text rows, the tile wants the 64
//               tokens of a clip together) and three 32-row boxes of W (this CTA's half of the 192 rows), 6-stage ring
//   warp 1      MMA issuer (leader CTA, one thread); two accumulator stages of 192 columns in TMEM
//   warps 2..9  epilogue, two per TMEM lane quadrant:
//       phase 1  tcgen05.ld of the quadrant's 32 rows (the warp with member = 0 takes dims 0..31 of Q, K and V, member = 1
//                dims 32..63) -> + bias -> bf16 -> 128B-swizzled shared tiles sQ / sK / sV [128 tokens][64]; accumulator
//                handed back to the MMA warp (the main loop of the next item but one starts)
//       phase 2  one 16-query m-tile per warp against the 64 keys of its clip: the mma.sync / ldmatrix flash step of
//                attention.cu (single chunk), additive mask, softmax, P V; O rows leave through the warp's (finished) 16 rows
//                of sQ as full 128-byte lines into ctx [tokens, H]
// The arithmetic is exactly that of gemm.cu's bf16 epilogue followed by attention.cu's kernel: results are bit-equal to the
// unfused pair of launches (tests/test_gpu_ops.py).  Inference forward only (no dropout, nothing saved for the backward).
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

namespace mra {
namespace {

constexpr int BM = 128;                       // token rows per CTA = 2 clips x (32 queries + 32 text tokens)
constexpr int BK = 64;
constexpr int BN = 192;                       // Q_h | K_h | V_h
constexpr int HD = 64;
constexpr int NQ = 32, NT = 32, S = NQ + NT;  // tokens of a clip
constexpr int STAGES = 6;
constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KiB
constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;      // 12 KiB: this CTA's half of the W slab
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int BOX = 32 * 128;                         // 32 rows x 64 bf16
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TILE_BYTES = BM * 128;                  // one of sQ / sK / sV: 128 tokens x 64 bf16
constexpr int QKV_OFFSET = STAGES * STAGE_BYTES;
constexpr int MASK_OFFSET = QKV_OFFSET + 3 * TILE_BYTES;          // [2 clips][64] fp32, log2 domain
constexpr int BAR_OFFSET = MASK_OFFSET + 2 * S * 4;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget exceeded");
constexpr uint32_t TMEM_COLS = 512;           // two accumulator stages of 192 columns at a stride of 256
constexpr int ACC_STRIDE = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_PROBLEMS = 2;

struct Maps {
    CUtensorMap x[MAX_PROBLEMS], w[MAX_PROBLEMS];
};
struct Params {
    const float* bias[MAX_PROBLEMS];
    const float* add_mask[MAX_PROBLEMS];
    __nv_bfloat16* ctx[MAX_PROBLEMS];
    int64_t ldo[MAX_PROBLEMS];
    int rows[MAX_PROBLEMS];          // clips
    int item_start[MAX_PROBLEMS + 1];   // first work item of each problem (items: clip-block major, head fastest)
    int problems, heads, K;
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// address of 16-byte unit `u` of token row `row` of a [tokens][64] bf16 tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int u) {
    return base + row * 128 + ((u ^ (row & 7)) << 4);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

__device__ __forceinline__ void decode_item(const Params& p, int item, int& g, int& blk, int& head) {
    g = (p.problems > 1 && item >= p.item_start[1]) ? 1 : 0;
    const int t = item - p.item_start[g];
    blk = t / p.heads;
    head = t - blk * p.heads;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
qkv_attn_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    const uint32_t sQ = ptx::smem_u32(smem + QKV_OFFSET);
    const uint32_t sK = sQ + TILE_BYTES, sV = sQ + 2 * TILE_BYTES;
    float* sMask = reinterpret_cast<float*>(smem + MASK_OFFSET);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = ptx::cluster_ctarank();
    const int first_item = static_cast<int>(blockIdx.x) / 2;
    const int item_stride = static_cast<int>(gridDim.x) / 2;
    const int total_items = p.item_start[p.problems];
    const int k_blocks = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < p.problems; ++g) {
            ptx::prefetch_tensormap(&maps.x[g]);
            ptx::prefetch_tensormap(&maps.w[g]);
        }
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull_bar[i], 1);
            ptx::mbar_init(&tempty_bar[i], 2 * EPI_WARPS);   // one arrival per epilogue warp of both CTAs (leader's barrier)
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    ptx::cluster_sync_all();   // the peer's barriers exist before anything is signalled to them
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
                int g, blk, head;
                decode_item(p, item, g, blk, head);
                const CUtensorMap* tmX = &maps.x[g];
                const CUtensorMap* tmW = &maps.w[g];
                const int H = p.heads * HD;
                const int clip0 = blk * 4 + static_cast<int>(crank) * 2;   // this CTA's two clips
                const int Mq = p.rows[g] * NQ;
                // W rows of this CTA's half of (Q_h | K_h | V_h): three boxes of 32 rows
                int wrow[3];
                if (crank == 0) { wrow[0] = head * HD; wrow[1] = head * HD + 32; wrow[2] = H + head * HD; }
                else { wrow[0] = H + head * HD + 32; wrow[1] = 2 * H + head * HD; wrow[2] = 2 * H + head * HD + 32; }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the pair
                    const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
                    if (crank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    uint8_t* a = sA + stage * A_STAGE_BYTES;
                    uint8_t* b = sB + stage * B_STAGE_BYTES;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        // rows past the end of the token matrix are zero-filled by the TMA unit (clip blocks at the tail)
                        ptx::tma_load_2d_pair(a + (2 * c) * BOX, tmX, lead_bar, kb * BK, (clip0 + c) * NQ);
                        ptx::tma_load_2d_pair(a + (2 * c + 1) * BOX, tmX, lead_bar, kb * BK, Mq + (clip0 + c) * NT);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) ptx::tma_load_2d_pair(b + j * BOX, tmW, lead_bar, kb * BK, wrow[j]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA, single thread)
        if (lane == 0 && crank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int item = first_item; item < total_items; item += item_stride, ++iter) {
                const int acc = iter & 1;
                ptx::mbar_wait(&tempty_bar[acc], ((iter >> 1) & 1) ^ 1);   // the epilogues of both CTAs have drained this stage
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA + stage * A_STAGE_BYTES));
                    const uint64_t b_desc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        ptx::umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    ptx::umma_commit_pair(&empty_bar[stage], static_cast<uint16_t>(0x3));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_pair(&tfull_bar[acc], static_cast<uint16_t>(0x3));
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..9
        const int quad = warp & 3;            // TMEM lanes / tile rows [32 quad, 32 quad + 32): clip quad / 2, queries (even) or text
        const int member = (warp - 2) >> 2;   // phase 1: dims [32 member, 32 member + 32) of Q, K and V
        const int clip_l = quad >> 1;         // this warp's clip inside the CTA
        const int mt = (quad & 1) * 2 + member;   // phase 2: m-tile (16 queries) of the clip: rows [16 mt, 16 mt + 16)
        const int g8 = lane >> 2, t4 = lane & 3;
        const float scale_log2 = 0.125f * LOG2E;
        const int trow = quad * 32 + lane;    // phase 1: this thread's token row of the tile
        int iter = 0;
        for (int item = first_item; item < total_items; item += item_stride, ++iter) {
            int g, blk, head;
            decode_item(p, item, g, blk, head);
            const int H = p.heads * HD;
            const int acc = iter & 1;
            const int rows = p.rows[g];
            const int clip = blk * 4 + static_cast<int>(crank) * 2 + clip_l;   // global clip of this warp
            // bias of this warp's 3 x 32 columns (lane l keeps column l of each unit), fetched before the accumulator wait
            float bq[3];
#pragma unroll
            for (int u = 0; u < 3; ++u)
                bq[u] = p.bias[g] != nullptr ? __ldg(p.bias[g] + u * H + head * HD + member * 32 + lane) : 0.f;
            // additive mask of this CTA's two clips (log2 domain): 128 values, written by the first four warps
            if (warp - 2 < 4) {
                const int j = (warp - 2) * 32 + lane;        // [clip_l'][key]
                const int c = blk * 4 + static_cast<int>(crank) * 2 + (j >> 6);
                float mv = 0.f;
                if (p.add_mask[g] != nullptr && c < rows) mv = __ldg(p.add_mask[g] + static_cast<int64_t>(c) * S + (j & 63)) * LOG2E;
                sMask[j] = mv;
            }
            ptx::mbar_wait(&tfull_bar[acc], (iter >> 1) & 1);
            ptx::tc_fence_after();
            // ---- phase 1: accumulator -> + bias -> bf16 -> sQ / sK / sV
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * ACC_STRIDE + member * 32;
            uint32_t racc[3][32];
#pragma unroll
            for (int u = 0; u < 3; ++u) ptx::tmem_ld_32x32b_x32(t_row + u * HD, racc[u]);
            ptx::tmem_ld_wait();
            // last TMEM read of this item by this warp: the MMA warp may reuse the accumulator stage
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[acc]), 0));
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const uint32_t dst = sQ + u * TILE_BYTES;
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // 16-byte unit member * 4 + q of the row: dims 32 member + 8 q ..
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 8 * q + 2 * e;
                        const float v0 = __uint_as_float(racc[u][j]) + __shfl_sync(0xffffffffu, bq[u], j);
                        const float v1 = __uint_as_float(racc[u][j + 1]) + __shfl_sync(0xffffffffu, bq[u], j + 1);
                        w[e] = ptx::pack_bf16x2(v0, v1);
                    }
                    ptx::st_shared_v4(tile_addr(dst, trow, member * 4 + q), w[0], w[1], w[2], w[3]);
                }
            }
            epi_bar_sync();   // Q / K / V tiles and the mask of both clips are complete
            // ---- phase 2: 16 queries x 64 keys of this warp's clip (attention.cu's single-chunk step)
            {
                const int qr0 = clip_l * S + mt * 16;          // first query row of the m-tile inside the tile
                const int kr0 = clip_l * S;                    // first key row of the clip
                const float* mk = sMask + clip_l * S;
                uint32_t qf[4][4];
                {
                    const int row = qr0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], tile_addr(sQ, row, ks * 2 + (lane >> 4)));
                }
                float s[8][4];
#pragma unroll
                for (int np = 0; np < 4; ++np) {    // pairs of 8-key tiles
                    s[2 * np][0] = s[2 * np][1] = s[2 * np][2] = s[2 * np][3] = 0.f;
                    s[2 * np + 1][0] = s[2 * np + 1][1] = s[2 * np + 1][2] = s[2 * np + 1][3] = 0.f;
                    const int key = kr0 + np * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t kf[4];
                        ldmatrix_x4(kf, tile_addr(sK, key, ks * 2 + ((lane >> 3) & 1)));
                        mma_bf16_16816(s[2 * np], qf[ks], kf[0], kf[1]);
                        mma_bf16_16816(s[2 * np + 1], qf[ks], kf[2], kf[3]);
                    }
                }
                float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float2 mm = *reinterpret_cast<const float2*>(mk + nt * 8 + 2 * t4);
                    s[nt][0] = fmaf(s[nt][0], scale_log2, mm.x);
                    s[nt][1] = fmaf(s[nt][1], scale_log2, mm.y);
                    s[nt][2] = fmaf(s[nt][2], scale_log2, mm.x);
                    s[nt][3] = fmaf(s[nt][3], scale_log2, mm.y);
                    mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                    mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
                    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
                }
                uint32_t pf[4][4];
                float ls[2] = {0.f, 0.f};
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float p0 = exp2f(s[nt][0] - mx[0]);
                    const float p1 = exp2f(s[nt][1] - mx[0]);
                    const float p2 = exp2f(s[nt][2] - mx[1]);
                    const float p3 = exp2f(s[nt][3] - mx[1]);
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
                float o_acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {          // 16 keys per step
                    const int key = kr0 + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
                    for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide dim tiles
                        uint32_t vf[4];
                        ldmatrix_x4_t(vf, tile_addr(sV, key, dp * 2 + (lane >> 4)));
                        mma_bf16_16816(o_acc[2 * dp], pf[j], vf[0], vf[1]);
                        mma_bf16_16816(o_acc[2 * dp + 1], pf[j], vf[2], vf[3]);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    ls[h] += __shfl_xor_sync(0xffffffffu, ls[h], 1);
                    ls[h] += __shfl_xor_sync(0xffffffffu, ls[h], 2);
                }
                const float inv[2] = {1.f / ls[0], 1.f / ls[1]};
                // O through this warp's own 16 rows of sQ (only this warp read them), then full 128-byte lines to ctx
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row = qr0 + g8 + h * 8;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        ptx::st_shared_b32(tile_addr(sQ, row, nt) + 4 * t4,
                                           ptx::pack_bf16x2(o_acc[nt][2 * h] * inv[h], o_acc[nt][2 * h + 1] * inv[h]));
                }
                __syncwarp();
                if (clip < rows) {
                    __nv_bfloat16* ctx = p.ctx[g];
                    const int64_t ldo = p.ldo[g];
                    const int64_t Mq = static_cast<int64_t>(rows) * NQ;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int lr = mt * 16 + i * 4 + (lane >> 3);          // token of the clip: < 32 query, else text
                        const uint4 val = ptx::ld_shared_v4(tile_addr(sQ, clip_l * S + lr, lane & 7));
                        const int64_t grow = lr < NQ ? static_cast<int64_t>(clip) * NQ + lr : Mq + static_cast<int64_t>(clip) * NT + (lr - NQ);
                        *reinterpret_cast<uint4*>(ctx + grow * ldo + head * HD + (lane & 7) * 8) = val;
                    }
                }
            }
            epi_bar_sync();   // every warp is done with the tiles: the next item's phase 1 may overwrite them
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();   // no CTA exits while its peer may still signal it / read its shared memory
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

}  // namespace

// n <= 2 problems (the video and the audio Q-Former of a lockstep forward) sharing heads and K
int launch_qkv_attention(const QkvAttnArgs* a, int n, cudaStream_t s) {
    MRA_REQUIRE(n >= 1 && n <= MAX_PROBLEMS, "fused QKV + attention takes 1..%d problems, got %d", MAX_PROBLEMS, n);
    Maps maps;
    Params p;
    p.problems = n;
    p.heads = a[0].heads;
    p.K = a[0].K;
    int total = 0;
    for (int g = 0; g < MAX_PROBLEMS; ++g) {
        const QkvAttnArgs& q = a[g < n ? g : 0];
        if (g < n) {
            MRA_REQUIRE(q.x && q.w && q.ctx, "fused QKV + attention: NULL operand");
            MRA_REQUIRE(q.rows > 0 && q.heads == p.heads && q.heads > 0 && q.K == p.K && q.K > 0 && q.K % 8 == 0,
                        "fused QKV + attention: problems must share heads and K (rows=%d heads=%d K=%d)", q.rows, q.heads, q.K);
            MRA_REQUIRE((reinterpret_cast<uintptr_t>(q.ctx) & 15) == 0 && q.ldo % 8 == 0 && q.ldo >= q.heads * HD,
                        "fused QKV + attention: ctx rows must be 16-byte aligned");
            const int64_t Mtot = static_cast<int64_t>(q.rows) * S;
            if (int e = get_tensor_map(q.x, Mtot, q.K, q.ldx, 32, BK, 2, &maps.x[g])) return e;
            if (int e = get_tensor_map(q.w, 3 * q.heads * HD, q.K, q.ldw, 32, BK, 2, &maps.w[g])) return e;
            p.bias[g] = q.bias; p.add_mask[g] = q.add_mask;
            p.ctx[g] = reinterpret_cast<__nv_bfloat16*>(q.ctx); p.ldo[g] = q.ldo;
            p.rows[g] = q.rows;
            p.item_start[g] = total;
            total += ((q.rows + 3) / 4) * q.heads;
        } else {
            maps.x[g] = maps.x[0]; maps.w[g] = maps.w[0];
            p.bias[g] = nullptr; p.add_mask[g] = nullptr; p.ctx[g] = nullptr; p.ldo[g] = 0; p.rows[g] = 0;
            p.item_start[g] = total;
        }
    }
    p.item_start[MAX_PROBLEMS] = total;
    for (int g = n; g <= MAX_PROBLEMS; ++g) p.item_start[g] = total;
    if (int e = ensure_smem_attr(reinterpret_cast<const void*>(qkv_attn_kernel), SMEM_TOTAL)) return e;
    const int pairs = total < sm_count() / 2 ? total : sm_count() / 2;
    // (the cluster size is the kernel's compile-time __cluster_dims__: no launch attribute for it)
    MRA_CHECK_CUDA(launch_pdl(qkv_attn_kernel, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_TOTAL, s, 1, maps, p));
    return 0;
}

}  // namespace mra
