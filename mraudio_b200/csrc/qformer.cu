// Q-Former forward orchestration: one call enqueues the whole `Qformer.bert(...)` + `llm_proj` of one modality on the
// caller's stream (models/xinstructblip.py:286-293 + :303).  All memory is caller-owned; the handle keeps only the
// configuration, the weight pointers and a launch counter.
//
// HBM layout of the activations ("split" layout): hidden states live as ONE [rows*Nq + rows*T, H] matrix with the
// query tokens of all rows first and the text tokens after them, so every Linear of the layer stack is a plain 2-D
// K-major GEMM (self-attention projections over the whole matrix, cross-attention / FFN_query over the first
// rows*Nq rows, FFN_text over the rest) and no torch.cat / slicing copies exist.  The residual stream is fp32
// (x32 / a32 / pre), each with a bf16 shadow that feeds the tensor cores (see DESIGN.md "precision").
#include <vector>

#include "common.h"

struct mra_qformer {
    mra_qformer_config cfg;
    mra_qformer_weights w;
    bool has_weights = false;
    int n_cross = 0;
    int cross_slot[MRA_MAX_LAYERS];  // index of the layer's K/V block inside w_ckv, or -1
    int last_launches = 0;
    int gemm_impl = MRA_GEMM_IMPL_TCGEN05;
    cudaEvent_t layer_done[MRA_MAX_LAYERS + 1] = {};   // optional: recorded by the backward when a layer's gradients are final
                                                       // (slot `layers`: the projection's)
    bool fuse_ln = true;   // Linear + residual + LayerNorm in one cluster kernel (MRA_NO_FUSED_LN=1 disables: A/B runs)
    bool split_res = true; // with fuse_ln: residual stream as a bf16 (hi, lo) pair instead of fp32 (MRA_SPLIT_RESIDUAL=0 disables)
    bool fuse_qkv_attn = true; // inference forward with 32 queries + 32 text tokens per row: QKV Linear + self-attention core in
                               // one kernel, qkv never reaches HBM (qkv_attn.cu; MRA_FUSE_QKV_ATTN=0 disables: A/B runs)
    bool head_major = true; // inference forward: QKV / cross-K/V GEMMs write [head][token][64] for the attention kernel's TMA
                            // boxes (MRA_HEAD_MAJOR=0 disables: A/B runs); the save-for-backward forward keeps [token][heads * 64]
    // optional per-category device timing (CUDA events on the caller's stream), see mra_qformer_profile_*
    int profile_mode = MRA_PROFILE_OFF;
    struct Span { int cat; cudaEvent_t a, b; };
    std::vector<Span> spans;          // recorded since the last read
    std::vector<cudaEvent_t> pool;    // recycled events
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
};

namespace mra {
namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// dropout parameters of a call (thr8 == 0: off), see dropout.cuh
inline DropoutParams dropout_base(const mra_qformer_io* io) {
    DropoutParams d;
    if (io->dropout_p > 0.f) {
        int thr = static_cast<int>(io->dropout_p * 256.0f + 0.5f);
        thr = thr < 1 ? 1 : (thr > 255 ? 255 : thr);
        d.thr8 = static_cast<uint32_t>(thr);
        d.scale = 256.0f / static_cast<float>(256 - thr);
        d.seed_lo = static_cast<uint32_t>(io->dropout_seed);
        d.seed_hi = static_cast<uint32_t>(io->dropout_seed >> 32);
    }
    return d;
}

// Per-layer activation buffers.  Inference: every layer aliases ONE set (activations are dead after the layer).
// MRA_FWD_SAVE_FOR_BACKWARD: every layer has its own set, kept for mra_qformer_backward.
struct LayerBufs {
    __nv_bfloat16* xb;      // [Mtot, H]   layer input (bf16 GEMM operand); slot `layers` = output of the last layer
    __nv_bfloat16* qkv;     // [Mtot, 3H]
    __nv_bfloat16* ctx;     // [Mtot, H]   self-attention context
    float* pre_a;           // [Mtot, H]   pre-LayerNorm sum of the self-attention block
    __nv_bfloat16* ab;      // [Mtot, H]   LN_a output
    __nv_bfloat16* cq;      // [Mq, H]     cross-attention queries
    __nv_bfloat16* cctx;    // [Mq, H]     cross-attention context
    float* pre_c;           // [Mq, H]
    __nv_bfloat16* ab2;     // [Mq, H]     LN_c output (== ab rows [0, Mq) when not saving)
    __nv_bfloat16* z;       // [Mtot, I]   pre-GELU activations (saved only)
    __nv_bfloat16* inter;   // [Mtot, I]   GELU output
    float* pre_f;           // [Mtot, H]   pre-LayerNorm sums of FFN_query (rows < Mq) / FFN_text
};

struct Workspace {
    float* x32;             // residual stream: layer input (fp32)
    float* a32;             // residual stream: attention output (fp32)
    // split residual stream (inference with the fused GEMM+LayerNorm): value = hi + lo with hi = the bf16 GEMM operand
    // (xb / ab) and lo = bf16(value - hi) kept here; these alias the first half of x32 / a32
    __nv_bfloat16* xlo;
    __nv_bfloat16* alo;
    float* pre_e;           // embedding sums before the embedding LayerNorm (saved only)
    __nv_bfloat16* kv;      // [rows*Nk, ncross*2H]
    float* self_mask;       // [rows, S]
    float* enc_mask;        // [rows, Nk]
    LayerBufs layer[MRA_MAX_LAYERS + 1];
    size_t total;
};

Workspace carve(const mra_qformer* h, int rows, int T, int Nk, uint32_t flags, void* base) {
    const auto& c = h->cfg;
    const bool save = (flags & MRA_FWD_SAVE_FOR_BACKWARD) != 0;
    const size_t H = c.hidden, I = c.inter;
    const size_t Mq = static_cast<size_t>(rows) * c.num_query, Mtot = Mq + static_cast<size_t>(rows) * T;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return reinterpret_cast<uint8_t*>(base) + o;
    };
    Workspace w;
    w.x32 = reinterpret_cast<float*>(take(Mtot * H * 4));
    w.a32 = reinterpret_cast<float*>(take(Mtot * H * 4));
    w.xlo = reinterpret_cast<__nv_bfloat16*>(w.x32);
    w.alo = reinterpret_cast<__nv_bfloat16*>(w.a32);
    w.pre_e = save ? reinterpret_cast<float*>(take(Mtot * H * 4)) : nullptr;
    w.kv = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(rows) * Nk * h->n_cross * 2 * H * 2));
    w.self_mask = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * (c.num_query + T) * 4));
    w.enc_mask = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * Nk * 4));
    const int nsets = save ? c.layers : 1;
    for (int l = 0; l < nsets; ++l) {
        LayerBufs& b = w.layer[l];
        b.xb = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
        b.qkv = reinterpret_cast<__nv_bfloat16*>(take(Mtot * 3 * H * 2));
        b.ctx = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
        b.pre_a = reinterpret_cast<float*>(take(Mtot * H * 4));
        b.ab = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
        b.cq = reinterpret_cast<__nv_bfloat16*>(take(Mq * H * 2));
        b.inter = reinterpret_cast<__nv_bfloat16*>(take(Mtot * I * 2));
        if (save) {
            b.cctx = reinterpret_cast<__nv_bfloat16*>(take(Mq * H * 2));
            b.pre_c = reinterpret_cast<float*>(take(Mq * H * 4));
            b.ab2 = reinterpret_cast<__nv_bfloat16*>(take(Mq * H * 2));
            b.z = reinterpret_cast<__nv_bfloat16*>(take(Mtot * I * 2));
            b.pre_f = reinterpret_cast<float*>(take(Mtot * H * 4));
        } else {
            b.cctx = b.ctx;
            b.pre_c = b.pre_a;
            b.ab2 = b.ab;
            b.z = nullptr;
            b.pre_f = b.pre_a;
        }
    }
    if (save) {
        w.layer[c.layers] = w.layer[0];
        w.layer[c.layers].xb = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    } else {
        for (int l = 1; l <= c.layers; ++l) w.layer[l] = w.layer[0];
    }
    w.total = off;
    return w;
}

// scratch of the backward pass (caller-owned, see mra_qformer_backward_workspace_bytes)
struct BwdWorkspace {
    float* g_x; float* g_a; float* g_pre32;      // [Mtot, H] fp32
    __nv_bfloat16* g_pre16; __nv_bfloat16* g_ctx16;   // [Mtot, H]
    __nv_bfloat16* g_cq16;                        // [Mq, H]
    __nv_bfloat16* g_big16; __nv_bfloat16* g_big2;    // [Mtot, max(I, 3H)]
    __nv_bfloat16* g_kv16;                        // [rows*Nk, ncross*2H]: dL/d(keys, values) of ALL cross layers (one wgrad at the end)
    size_t total;
};

BwdWorkspace carve_bwd(const mra_qformer* h, int rows, int T, int Nk, void* base) {
    const auto& c = h->cfg;
    const size_t H = c.hidden, I = c.inter;
    const size_t Mq = static_cast<size_t>(rows) * c.num_query, Mtot = Mq + static_cast<size_t>(rows) * T;
    const size_t big = I > 3 * H ? I : 3 * H;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 1024);
        return reinterpret_cast<uint8_t*>(base) + o;
    };
    BwdWorkspace b;
    b.g_x = reinterpret_cast<float*>(take(Mtot * H * 4));
    b.g_a = reinterpret_cast<float*>(take(Mtot * H * 4));
    b.g_pre32 = reinterpret_cast<float*>(take(Mtot * H * 4));
    b.g_pre16 = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    b.g_ctx16 = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    b.g_cq16 = reinterpret_cast<__nv_bfloat16*>(take(Mq * H * 2));
    b.g_big16 = reinterpret_cast<__nv_bfloat16*>(take(Mtot * big * 2));
    b.g_big2 = reinterpret_cast<__nv_bfloat16*>(take(Mtot * big * 2));
    b.g_kv16 = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(rows) * Nk * h->n_cross * 2 * H * 2));
    b.total = off;
    return b;
}

}  // namespace
}  // namespace mra

using namespace mra;

extern "C" int mra_qformer_create(const mra_qformer_config* cfg, mra_qformer_t** out) {
    MRA_REQUIRE(cfg && out, "mra_qformer_create: NULL argument");
    MRA_REQUIRE(cfg->layers >= 1 && cfg->layers <= MRA_MAX_LAYERS, "layers %d out of range [1, %d]", cfg->layers, MRA_MAX_LAYERS);
    MRA_REQUIRE(cfg->heads > 0 && cfg->hidden == cfg->heads * 64, "head_dim must be 64 (hidden %d, heads %d)", cfg->hidden, cfg->heads);
    MRA_REQUIRE(cfg->hidden % 8 == 0 && cfg->inter % 8 == 0 && cfg->enc_width % 8 == 0 && cfg->llm_dim % 8 == 0,
                "hidden / inter / enc_width / llm_dim must be multiples of 8");
    MRA_REQUIRE(cfg->cross_freq >= 1 && cfg->num_query >= 1, "cross_freq and num_query must be positive");
    mra_qformer* h = new mra_qformer();
    h->cfg = *cfg;
    h->n_cross = 0;
    for (int l = 0; l < cfg->layers; ++l) h->cross_slot[l] = (l % cfg->cross_freq == 0) ? h->n_cross++ : -1;
    const char* impl = getenv("MRA_GEMM_IMPL");
    if (impl && std::string(impl) == "simt") h->gemm_impl = MRA_GEMM_IMPL_SIMT_DEBUG;
    if (getenv("MRA_NO_FUSED_LN")) h->fuse_ln = false;
    if (const char* e = getenv("MRA_SPLIT_RESIDUAL")) h->split_res = atoi(e) != 0;
    if (const char* e = getenv("MRA_HEAD_MAJOR")) h->head_major = atoi(e) != 0;
    if (const char* e = getenv("MRA_FUSE_QKV_ATTN")) h->fuse_qkv_attn = atoi(e) != 0;
    *out = h;
    return 0;
}

extern "C" int mra_qformer_set_weights(mra_qformer_t* h, const mra_qformer_weights* w) {
    MRA_REQUIRE(h && w, "mra_qformer_set_weights: NULL argument");
    MRA_REQUIRE(w->ln_e_g && w->ln_e_b && w->w_ckv && w->b_ckv, "embedding LayerNorm / cross K,V weights missing");
    for (int l = 0; l < h->cfg.layers; ++l) {
        const auto& L = w->layer[l];
        MRA_REQUIRE(L.w_qkv && L.b_qkv && L.w_ao && L.b_ao && L.ln_a_g && L.ln_a_b && L.w_fq1 && L.b_fq1 && L.w_fq2 &&
                        L.b_fq2 && L.ln_fq_g && L.ln_fq_b,
                    "layer %d: self-attention / query-FFN weights missing", l);
        if (h->cross_slot[l] >= 0)
            MRA_REQUIRE(L.w_cq && L.b_cq && L.w_co && L.b_co && L.ln_c_g && L.ln_c_b, "layer %d: cross-attention weights missing", l);
    }
    h->w = *w;
    h->has_weights = true;
    return 0;
}

extern "C" void mra_qformer_destroy(mra_qformer_t* h) {
    if (!h) return;
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : h->pool) cudaEventDestroy(e);
    delete h;
}

extern "C" size_t mra_qformer_workspace_bytes(const mra_qformer_t* h, int32_t rows, int32_t T, int32_t Nk, uint32_t flags) {
    if (!h || rows <= 0 || T < 0 || Nk <= 0) return 0;
    return carve(h, rows, T, Nk, flags, nullptr).total;
}

extern "C" size_t mra_qformer_backward_workspace_bytes(const mra_qformer_t* h, int32_t rows, int32_t T, int32_t Nk) {
    if (!h || rows <= 0 || T < 0 || Nk <= 0) return 0;
    return carve_bwd(h, rows, T, Nk, nullptr).total;
}

extern "C" int mra_qformer_profile_mode(mra_qformer_t* h, int32_t mode) {
    MRA_REQUIRE(h, "mra_qformer_profile_mode: NULL handle");
    MRA_REQUIRE(mode >= MRA_PROFILE_OFF && mode <= MRA_PROFILE_ALL, "unknown profile mode %d", mode);
    h->profile_mode = mode;
    return 0;
}

extern "C" int mra_qformer_profile_read(mra_qformer_t* h, double* ms_by_cat, int64_t* launches_by_cat) {
    MRA_REQUIRE(h && ms_by_cat && launches_by_cat, "mra_qformer_profile_read: NULL argument");
    for (int i = 0; i < MRA_NUM_CATS; ++i) { ms_by_cat[i] = 0.0; launches_by_cat[i] = 0; }
    for (auto& sp : h->spans) {
        MRA_CHECK_CUDA(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        MRA_CHECK_CUDA(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms_by_cat[sp.cat] += ms;
        launches_by_cat[sp.cat] += 1;
        h->pool.push_back(sp.a);
        h->pool.push_back(sp.b);
    }
    h->spans.clear();
    return 0;
}

extern "C" int mra_qformer_last_launch_count(const mra_qformer_t* h) { return h ? h->last_launches : 0; }

// Forward of up to two Q-Formers in lockstep (e.g. the video and the audio Q-Former of one batch): every Linear of a layer
// is ONE grouped GEMM launch over all contexts -- and FFN_query / FFN_text are further problems of the same launch -- so
// the persistent GEMM grid sees 2-4x the tiles (smaller wave-quantisation tails) at half the launch count.  The
// contexts must agree on the layer geometry (hidden, heads, layers, intermediate, queries, cross frequency, llm_dim);
// encoder width, rows, T and Nk may differ.
namespace {

constexpr int MAX_CTX = 2;

struct Ctx {
    mra_qformer* h;
    const mra_qformer_io* io;
    Workspace ws;
    int rows, T, Nk, S, Mq, Mt, Mtot, kv_ld;
    bool save;
    const float* self_mask;
    const float* enc_mask;
};

int forward_multi(int n, mra_qformer_t* const* hs, const mra_qformer_io* const* ios, void* const* workspaces,
                  const size_t* workspace_bytes, cudaStream_t s) {
    MRA_REQUIRE(n >= 1 && n <= MAX_CTX, "forward takes 1..%d Q-Formers per call, got %d", MAX_CTX, n);
    if (int e = device_check()) return e;
    Ctx cx[MAX_CTX];
    mra_qformer* h0 = hs[0];
    const auto& c = h0->cfg;
    const int Nq = c.num_query, H = c.hidden, I = c.inter;
    for (int i = 0; i < n; ++i) {
        mra_qformer* h = hs[i];
        const mra_qformer_io* io = ios[i];
        MRA_REQUIRE(h && io && workspaces[i], "mra_qformer_forward: NULL argument");
        MRA_REQUIRE(h->has_weights, "mra_qformer_forward: weights not set");
        const auto& ci = h->cfg;
        MRA_REQUIRE(ci.hidden == c.hidden && ci.layers == c.layers && ci.heads == c.heads && ci.inter == c.inter &&
                        ci.cross_freq == c.cross_freq && ci.num_query == c.num_query && ci.llm_dim == c.llm_dim &&
                        ci.ln_eps == c.ln_eps,
                    "Q-Formers run in lockstep must share the layer geometry");
        const auto& W = h->w;
        const int rows = io->rows, T = io->T, Nk = io->Nk;
        MRA_REQUIRE(rows > 0 && Nk > 0 && T >= 0, "bad shape rows=%d T=%d Nk=%d", rows, T, Nk);
        MRA_REQUIRE(T <= ci.max_pos, "T=%d exceeds max_position_embeddings=%d", T, ci.max_pos);
        MRA_REQUIRE(io->enc && io->query_embeds, "enc / query_embeds must not be NULL");
        MRA_REQUIRE(io->q_rows == 1 || io->q_rows == rows, "query_embeds rows must be 1 or %d, got %d", rows, io->q_rows);
        MRA_REQUIRE(T == 0 || (io->input_ids && W.word_emb && W.pos_emb), "text tokens need input_ids and embedding tables");
        MRA_REQUIRE(T == 0 || W.layer[0].w_ft1, "text tokens need the text FFN weights");
        MRA_REQUIRE(!io->llm_out || (W.w_proj && W.b_proj && ci.llm_dim > 0), "llm_out requested but no projection weights");
        MRA_REQUIRE((io->flags & MRA_FWD_SAVE_FOR_BACKWARD) == (ios[0]->flags & MRA_FWD_SAVE_FOR_BACKWARD) &&
                        (io->llm_out != nullptr) == (ios[0]->llm_out != nullptr),
                    "Q-Formers run in lockstep must agree on SAVE_FOR_BACKWARD and on llm_out");
        Ctx& x = cx[i];
        x.h = h; x.io = io;
        x.save = (io->flags & MRA_FWD_SAVE_FOR_BACKWARD) != 0;
        x.ws = carve(h, rows, T, Nk, io->flags, workspaces[i]);
        MRA_REQUIRE(workspace_bytes[i] >= x.ws.total, "workspace too small: %zu < %zu bytes", workspace_bytes[i], x.ws.total);
        MRA_REQUIRE((reinterpret_cast<uintptr_t>(workspaces[i]) & 255) == 0, "workspace must be 256-byte aligned");
        x.rows = rows; x.T = T; x.Nk = Nk; x.S = Nq + T;
        x.Mq = rows * Nq; x.Mt = rows * T; x.Mtot = x.Mq + x.Mt;
        x.kv_ld = h->n_cross * 2 * H;
        x.self_mask = x.enc_mask = nullptr;
    }
    const bool save = cx[0].save;
    const DropoutParams drop0 = dropout_base(ios[0]);
    if (drop0.thr8 != 0) {
        MRA_REQUIRE(n == 1 && save, "dropout_p > 0 needs MRA_FWD_SAVE_FOR_BACKWARD and one Q-Former per call (the training forward)");
        MRA_REQUIRE(ios[0]->dropout_p < 1.f && ios[0]->T % 32 == 0, "dropout: p must be < 1 and the text length a multiple of 32 (T=%d)", ios[0]->T);
    }
    auto dsite = [&](int site, int layer) { return drop0.thr8 != 0 ? dropout_site(drop0, site, layer) : DropoutParams(); };
    const DropoutParams drop_emb = dsite(DROP_EMB, 0);
    int launches = 0;
    // profiling spans (recorded on the first handle): MRA_PROFILE_DOMINANT brackets only GEMM launches, _ALL every launch
    int cur_cat = -1;
    cudaEvent_t cur_ev = nullptr;
    auto span_begin = [&](int cat) {
        cur_cat = -1;
        if (h0->profile_mode == MRA_PROFILE_OFF) return;
        if (h0->profile_mode == MRA_PROFILE_DOMINANT && cat > MRA_CAT_GEMM) return;
        cur_cat = cat;
        cur_ev = h0->get_event();
        cudaEventRecord(cur_ev, s);
    };
    auto span_end = [&]() {
        if (cur_cat < 0) return;
        cudaEvent_t b = h0->get_event();
        cudaEventRecord(b, s);
        h0->spans.push_back({cur_cat, cur_ev, b});
        cur_cat = -1;
    };
    // grouped GEMM launch over the problems queued with add()
    GemmArgs ga[4];
    int ng = 0;
    auto add = [&](const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, const float* res, int64_t ldr,
                   void* C, int64_t ldc, int M, int N, int K, int gelu, int f32) {
        if (M > 0) ga[ng++] = GemmArgs{A, lda, Wt, ldw, bias, res, ldr, C, ldc, M, N, K, gelu, f32};
    };
    auto flush = [&](int cat) -> int {
        if (ng == 0) return 0;
        int e = 0;
        span_begin(cat);
        if (h0->gemm_impl == MRA_GEMM_IMPL_SIMT_DEBUG) {
            for (int g = 0; g < ng && e == 0; ++g) { e = launch_gemm_simt(ga[g], s); ++launches; }
        } else {
            e = launch_gemm_tc_grouped(ga, ng, s);
            ++launches;
        }
        span_end();
        ng = 0;
        return e;
    };
    // Linear + residual + LayerNorm in one kernel (gemm_ln.cu) whenever the pre-LayerNorm sums need not be kept
    const bool fuse_ln = !save && H == 768 && h0->gemm_impl == MRA_GEMM_IMPL_TCGEN05 && h0->fuse_ln;
    // split residual stream: every post-LayerNorm tensor is a bf16 (hi, lo) pair, hi doubling as the next GEMM operand
    const bool split = fuse_ln && h0->split_res;
    // head-major Q / K / V: only the attention kernel reads these tensors in the inference forward (the backward's kernels
    // read the [token][heads * 64] form, so the save-for-backward forward keeps it)
    const bool hm = !save && c.hidden == c.heads * 64 && h0->gemm_impl == MRA_GEMM_IMPL_TCGEN05 && h0->head_major;
    // QKV Linear + self-attention core in one kernel when every context has the 32 + 32 token geometry
    bool fuse_qa = !save && c.hidden == c.heads * 64 && Nq == 32 && h0->gemm_impl == MRA_GEMM_IMPL_TCGEN05 && h0->fuse_qkv_attn;
    for (int i = 0; i < n; ++i) fuse_qa = fuse_qa && cx[i].T == 32;
    GemmLnArgs gl[4];
    int ngl = 0;
    // residual = res32 (fp32) or, in split form, res_hi + res_lo; outputs y32 + y16 or y16 (hi) + y_lo
    auto add_gl = [&](const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, const float* res32,
                      const void* res_hi, const void* res_lo, const float* g, const float* b, float* y32, void* y16, void* y_lo,
                      int M, int K) {
        if (M <= 0) return;
        GemmLnArgs a{A, lda, Wt, ldw, bias, split ? res_hi : static_cast<const void*>(res32), H, g, b, split ? nullptr : y32, H, y16, H, M, K};
        if (split) { a.res_lo = res_lo; a.y_lo = y_lo; }
        gl[ngl++] = a;
    };
    auto flush_gl = [&]() -> int {
        if (ngl == 0) return 0;
        span_begin(MRA_CAT_GEMM);
        int e = launch_gemm_ln_grouped(gl, ngl, c.ln_eps, s);
        span_end();
        ++launches;
        ngl = 0;
        return e;
    };
    LnSegment ls[4];
    int nl = 0;
    auto add_ln = [&](const float* pre, const float* g, const float* b, float* y32, void* y16, int M) {
        if (M > 0) ls[nl++] = LnSegment{pre, g, b, y32, y16, M};
    };
    auto flush_ln = [&]() -> int {
        if (nl == 0) return 0;
        span_begin(MRA_CAT_LAYERNORM);
        int e = launch_layernorm_grouped(ls, nl, H, c.ln_eps, s);
        span_end();
        ++launches;
        nl = 0;
        return e;
    };
#define MRA_TRY(expr)            \
    do {                         \
        if (int _e = (expr)) return _e; \
    } while (0)

    // ---- embeddings + masks + cross-attention keys / values of ALL cross layers (one GEMM per context: the encoder
    //      tokens are read once; the contexts differ in K = encoder width, so these are separate launches)
    for (int i = 0; i < n; ++i) {
        Ctx& x = cx[i];
        const auto& W = x.h->w;
        span_begin(MRA_CAT_OTHER);
        MRA_TRY(launch_embed_layernorm(x.io->query_embeds, x.io->q_rows, x.io->input_ids, W.word_emb, W.pos_emb, W.ln_e_g,
                                       W.ln_e_b, split ? nullptr : x.ws.x32, x.ws.layer[0].xb, split ? x.ws.xlo : nullptr,
                                       x.ws.pre_e, x.rows, Nq, x.T, H, x.h->cfg.vocab, c.ln_eps, s,
                                       drop0.thr8 != 0 ? &drop_emb : nullptr));
        span_end();
        ++launches;
        if (x.io->attn_mask) {
            MRA_TRY(launch_build_enc_mask(x.io->attn_mask, x.ws.self_mask, x.rows, x.S, s));
            ++launches;
            x.self_mask = x.ws.self_mask;
        }
        if (x.io->enc_mask) {
            MRA_TRY(launch_build_enc_mask(x.io->enc_mask, x.ws.enc_mask, x.rows, x.Nk, s));
            ++launches;
            x.enc_mask = x.ws.enc_mask;
        }
        const int Wd = x.h->cfg.enc_width;
        // first reader of the encoder tokens: they may still be on their way (mra_qformer_io::enc_ready)
        if (x.io->enc_ready) MRA_CHECK_CUDA(cudaStreamWaitEvent(s, reinterpret_cast<cudaEvent_t>(x.io->enc_ready), 0));
        add(x.io->enc, Wd, W.w_ckv, Wd, W.b_ckv, nullptr, 0, x.ws.kv, x.kv_ld, x.rows * x.Nk, x.kv_ld, Wd, 0, 0);
        if (hm && ng > 0) ga[ng - 1].c_head_major = 1;
        MRA_TRY(flush(MRA_CAT_GEMM_CROSS_KV));
    }

    for (int l = 0; l < c.layers; ++l) {
        const bool last = l == c.layers - 1;
        const bool cross = h0->cross_slot[l] >= 0;
        // ---- self-attention over queries || text
        if (fuse_qa) {
            QkvAttnArgs qa[MAX_CTX];
            for (int i = 0; i < n; ++i) {
                const auto& L = cx[i].h->w.layer[l];
                const LayerBufs& B = cx[i].ws.layer[l];
                qa[i] = QkvAttnArgs{B.xb, H, L.w_qkv, H, L.b_qkv, B.ctx, H, cx[i].self_mask, cx[i].rows, c.heads, H};
            }
            span_begin(MRA_CAT_GEMM);
            MRA_TRY(launch_qkv_attention(qa, n, s));
            span_end();
            ++launches;
        } else {
            for (int i = 0; i < n; ++i) {
                const auto& L = cx[i].h->w.layer[l];
                const LayerBufs& B = cx[i].ws.layer[l];
                add(B.xb, H, L.w_qkv, H, L.b_qkv, nullptr, 0, B.qkv, 3 * H, cx[i].Mtot, 3 * H, H, 0, 0);
                if (hm && ng > 0) ga[ng - 1].c_head_major = 1;
            }
            MRA_TRY(flush(MRA_CAT_GEMM));
            {
                AttnArgs aa[MAX_CTX];
                for (int i = 0; i < n; ++i) {
                    const LayerBufs& B = cx[i].ws.layer[l];
                    aa[i] = AttnArgs{B.qkv, 3 * H, B.qkv + H, 3 * H, B.qkv + 2 * H, 3 * H, B.ctx, H, cx[i].self_mask, cx[i].rows, c.heads,
                                     cx[i].S, cx[i].S, Nq, 0};
                    if (hm) {   // qkv = [3 * heads][Mtot][64]: Q heads, then K heads, then V heads
                        const int64_t hs = static_cast<int64_t>(cx[i].Mtot) * 64;
                        aa[i].q = B.qkv; aa[i].k = B.qkv + c.heads * hs; aa[i].v = B.qkv + 2 * c.heads * hs;
                        aa[i].ldq = aa[i].ldk = aa[i].ldv = 64;
                        aa[i].hsq = aa[i].hsk = aa[i].hsv = hs;
                    }
                    aa[i].drop = dsite(DROP_SELF_PROBS, l);
                }
                span_begin(MRA_CAT_ATTENTION);
                MRA_TRY(launch_attention_pair(aa, n, s, &launches));
                span_end();
            }
        }
        for (int i = 0; i < n; ++i) {
            const auto& L = cx[i].h->w.layer[l];
            const LayerBufs& B = cx[i].ws.layer[l];
            if (fuse_ln) {
                add_gl(B.ctx, H, L.w_ao, H, L.b_ao, cx[i].ws.x32, B.xb, cx[i].ws.xlo, L.ln_a_g, L.ln_a_b, cx[i].ws.a32, B.ab,
                       cx[i].ws.alo, cx[i].Mtot, H);
            } else {
                add(B.ctx, H, L.w_ao, H, L.b_ao, cx[i].ws.x32, H, B.pre_a, H, cx[i].Mtot, H, H, 0, 1);
                if (ng > 0) ga[ng - 1].drop = dsite(DROP_SELF_OUT, l);
                add_ln(B.pre_a, L.ln_a_g, L.ln_a_b, cx[i].ws.a32, B.ab, cx[i].Mtot);
            }
        }
        MRA_TRY(flush_gl());
        MRA_TRY(flush(MRA_CAT_GEMM));
        MRA_TRY(flush_ln());
        // ---- cross-attention of the query tokens onto this row's encoder tokens
        if (cross) {
            for (int i = 0; i < n; ++i) {
                const auto& L = cx[i].h->w.layer[l];
                const LayerBufs& B = cx[i].ws.layer[l];
                add(B.ab, H, L.w_cq, H, L.b_cq, nullptr, 0, B.cq, H, cx[i].Mq, H, H, 0, 0);
            }
            MRA_TRY(flush(MRA_CAT_GEMM));
            {
                AttnArgs aa[MAX_CTX];
                for (int i = 0; i < n; ++i) {
                    const LayerBufs& B = cx[i].ws.layer[l];
                    const __nv_bfloat16* kbase = cx[i].ws.kv + static_cast<size_t>(cx[i].h->cross_slot[l]) * 2 * H;
                    aa[i] = AttnArgs{B.cq, H, kbase, cx[i].kv_ld, kbase + H, cx[i].kv_ld, B.cctx, H, cx[i].enc_mask, cx[i].rows,
                                     c.heads, Nq, cx[i].Nk, Nq, 1};
                    if (hm) {   // kv = [cross layers * 2 * heads][rows * Nk][64]: per cross layer K heads, then V heads
                        const int64_t hs = static_cast<int64_t>(cx[i].rows) * cx[i].Nk * 64;
                        aa[i].k = cx[i].ws.kv + static_cast<int64_t>(cx[i].h->cross_slot[l]) * 2 * c.heads * hs;
                        aa[i].v = reinterpret_cast<const __nv_bfloat16*>(aa[i].k) + c.heads * hs;
                        aa[i].ldk = aa[i].ldv = 64;
                        aa[i].hsk = aa[i].hsv = hs;
                    }
                    aa[i].drop = dsite(DROP_CROSS_PROBS, l);
                }
                span_begin(MRA_CAT_ATTENTION);
                MRA_TRY(launch_attention_pair(aa, n, s, &launches));
                span_end();
            }
            for (int i = 0; i < n; ++i) {
                const auto& L = cx[i].h->w.layer[l];
                const LayerBufs& B = cx[i].ws.layer[l];
                if (fuse_ln) {
                    add_gl(B.cctx, H, L.w_co, H, L.b_co, cx[i].ws.a32, B.ab, cx[i].ws.alo, L.ln_c_g, L.ln_c_b, cx[i].ws.a32, B.ab2,
                           cx[i].ws.alo, cx[i].Mq, H);
                } else {
                    add(B.cctx, H, L.w_co, H, L.b_co, cx[i].ws.a32, H, B.pre_c, H, cx[i].Mq, H, H, 0, 1);
                    if (ng > 0) ga[ng - 1].drop = dsite(DROP_CROSS_OUT, l);
                    add_ln(B.pre_c, L.ln_c_g, L.ln_c_b, cx[i].ws.a32, B.ab2, cx[i].Mq);
                }
            }
            MRA_TRY(flush_gl());
            MRA_TRY(flush(MRA_CAT_GEMM));
            MRA_TRY(flush_ln());
        }
        // ---- FFN_query (rows < Mq) and FFN_text (the rest) of every context: ONE grouped launch per Linear
        bool text_on[MAX_CTX];
        for (int i = 0; i < n; ++i) {
            text_on[i] = cx[i].Mt > 0 && !(last && (cx[i].io->flags & MRA_FWD_SKIP_DEAD_TEXT_FFN));
            if (text_on[i]) {
                const auto& L = cx[i].h->w.layer[l];
                MRA_REQUIRE(L.w_ft1 && L.b_ft1 && L.w_ft2 && L.b_ft2 && L.ln_ft_g && L.ln_ft_b, "layer %d: text FFN weights missing", l);
            }
        }
        for (int i = 0; i < n; ++i) {
            const auto& L = cx[i].h->w.layer[l];
            const LayerBufs& B = cx[i].ws.layer[l];
            const __nv_bfloat16* fq_in = cross ? B.ab2 : B.ab;
            __nv_bfloat16* h1 = save ? B.z : B.inter;
            const size_t o = static_cast<size_t>(cx[i].Mq) * H, oi = static_cast<size_t>(cx[i].Mq) * I;
            add(fq_in, H, L.w_fq1, H, L.b_fq1, nullptr, 0, h1, I, cx[i].Mq, I, H, save ? 0 : 1, 0);
            if (text_on[i]) add(B.ab + o, H, L.w_ft1, H, L.b_ft1, nullptr, 0, h1 + oi, I, cx[i].Mt, I, H, save ? 0 : 1, 0);
        }
        MRA_TRY(flush(MRA_CAT_GEMM));
        if (save) {
            for (int i = 0; i < n; ++i) {
                const LayerBufs& B = cx[i].ws.layer[l];
                const int nrow = cx[i].Mq + (text_on[i] ? cx[i].Mt : 0);
                span_begin(MRA_CAT_OTHER);
                MRA_TRY(launch_gelu_fwd(B.z, B.inter, static_cast<int64_t>(nrow) * I, s));
                span_end();
                ++launches;
            }
        }
        for (int i = 0; i < n; ++i) {
            const auto& L = cx[i].h->w.layer[l];
            const LayerBufs& B = cx[i].ws.layer[l];
            __nv_bfloat16* xb_next = cx[i].ws.layer[l + 1].xb;
            const size_t o = static_cast<size_t>(cx[i].Mq) * H, oi = static_cast<size_t>(cx[i].Mq) * I;
            if (fuse_ln) {
                add_gl(B.inter, I, L.w_fq2, I, L.b_fq2, cx[i].ws.a32, cross ? B.ab2 : B.ab, cx[i].ws.alo, L.ln_fq_g, L.ln_fq_b,
                       cx[i].ws.x32, xb_next, cx[i].ws.xlo, cx[i].Mq, I);
                if (text_on[i])
                    add_gl(B.inter + oi, I, L.w_ft2, I, L.b_ft2, cx[i].ws.a32 + o, B.ab + o, cx[i].ws.alo + o, L.ln_ft_g, L.ln_ft_b,
                           cx[i].ws.x32 + o, xb_next + o, cx[i].ws.xlo + o, cx[i].Mt, I);
            } else {
                add(B.inter, I, L.w_fq2, I, L.b_fq2, cx[i].ws.a32, H, B.pre_f, H, cx[i].Mq, H, I, 0, 1);
                if (ng > 0) ga[ng - 1].drop = dsite(DROP_FFN_OUT, l);
                add_ln(B.pre_f, L.ln_fq_g, L.ln_fq_b, cx[i].ws.x32, xb_next, cx[i].Mq);
                if (text_on[i]) {
                    add(B.inter + oi, I, L.w_ft2, I, L.b_ft2, cx[i].ws.a32 + o, H, B.pre_f + o, H, cx[i].Mt, H, I, 0, 1);
                    if (ng > 0) { ga[ng - 1].drop = dsite(DROP_FFN_OUT, l); ga[ng - 1].drop_row0 = cx[i].Mq; }   // text rows follow the queries
                    add_ln(B.pre_f + o, L.ln_ft_g, L.ln_ft_b, cx[i].ws.x32 + o, xb_next + o, cx[i].Mt);
                }
            }
        }
        MRA_TRY(flush_gl());
        MRA_TRY(flush(MRA_CAT_GEMM));
        MRA_TRY(flush_ln());
        for (int i = 0; i < n; ++i) {
            if (cx[i].Mt > 0 && !text_on[i] && cx[i].io->last_hidden) {
                // dead last-layer text FFN: keep last_hidden well-defined by passing the attention output through
                const size_t o = static_cast<size_t>(cx[i].Mq) * H;
                if (split) {
                    MRA_CHECK_CUDA(cudaMemcpyAsync(cx[i].ws.layer[l + 1].xb + o, cx[i].ws.layer[l].ab + o,
                                                   static_cast<size_t>(cx[i].Mt) * H * 2, cudaMemcpyDeviceToDevice, s));
                    MRA_CHECK_CUDA(cudaMemcpyAsync(cx[i].ws.xlo + o, cx[i].ws.alo + o, static_cast<size_t>(cx[i].Mt) * H * 2,
                                                   cudaMemcpyDeviceToDevice, s));
                    ++launches;
                } else {
                    MRA_CHECK_CUDA(cudaMemcpyAsync(cx[i].ws.x32 + o, cx[i].ws.a32 + o, static_cast<size_t>(cx[i].Mt) * H * 4,
                                                   cudaMemcpyDeviceToDevice, s));
                }
                ++launches;
            }
        }
    }
    // ---- outputs
    for (int i = 0; i < n; ++i) {
        if (cx[i].io->last_hidden) {
            span_begin(MRA_CAT_OTHER);
            if (split) {
                MRA_TRY(launch_gather_last_hidden_split(cx[i].ws.layer[c.layers].xb, cx[i].ws.xlo, cx[i].io->last_hidden, cx[i].rows,
                                                        Nq, cx[i].T, H, s));
            } else {
                MRA_TRY(launch_gather_last_hidden(cx[i].ws.x32, cx[i].io->last_hidden, cx[i].rows, Nq, cx[i].T, H, s));
            }
            span_end();
            ++launches;
        }
        if (cx[i].io->llm_out) {
            const auto& W = cx[i].h->w;
            const mra_qformer_io* io = cx[i].io;
            add(cx[i].ws.layer[c.layers].xb, H, W.w_proj, H, W.b_proj, nullptr, 0, io->llm_out,
                io->llm_frames > 0 && io->llm_ld > 0 ? io->llm_ld : c.llm_dim, cx[i].Mq, c.llm_dim, H, 0, 0);
            if (io->llm_frames > 0) {   // scatter into the interleaved LLM prompt (4-D TMA store map)
                MRA_REQUIRE(Nq == 32 && cx[i].rows % io->llm_frames == 0 && h0->gemm_impl != MRA_GEMM_IMPL_SIMT_DEBUG,
                            "llm_frames=%d needs 32 queries and rows (%d) = videos * frames", io->llm_frames, cx[i].rows);
                GemmArgs& a = ga[ng - 1];
                a.c_frames = io->llm_frames;
                a.c_frame_stride = io->llm_frame_stride;
                a.c_batch_stride = io->llm_video_stride;
            }
        }
    }
    MRA_TRY(flush(MRA_CAT_GEMM));
#undef MRA_TRY
    for (int i = 0; i < n; ++i) hs[i]->last_launches = i == 0 ? launches : 0;
    return 0;
}

}  // namespace

extern "C" int mra_qformer_forward(mra_qformer_t* h, const mra_qformer_io* io, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
    MRA_REQUIRE(h && io && workspace, "mra_qformer_forward: NULL argument");
    return forward_multi(1, &h, &io, &workspace, &workspace_bytes, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int mra_qformer_forward_multi(int32_t n, mra_qformer_t* const* hs, const mra_qformer_io* const* ios,
                                         void* const* workspaces, const size_t* workspace_bytes, void* stream_) {
    MRA_REQUIRE(hs && ios && workspaces && workspace_bytes, "mra_qformer_forward_multi: NULL argument");
    return forward_multi(n, hs, ios, workspaces, workspace_bytes, reinterpret_cast<cudaStream_t>(stream_));
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of mra_qformer_forward(flags | MRA_FWD_SAVE_FOR_BACKWARD) for the Q-Former / projection parameters
// (encoders frozen: no gradient flows into `enc`).  Gradients are ACCUMULATED into `g` (zero them for a fresh step).
extern "C" int mra_qformer_backward(mra_qformer_t* h, const mra_qformer_io* io, const void* d_llm,
                                    const void* reserved, const mra_qformer_grads* g, void* workspace,
                                    size_t workspace_bytes, void* bwd_workspace, size_t bwd_bytes, void* stream_) {
    MRA_REQUIRE(h && io && d_llm && g && workspace && bwd_workspace, "mra_qformer_backward: NULL argument");
    MRA_REQUIRE(h->has_weights, "mra_qformer_backward: weights not set");
    MRA_REQUIRE(io->flags & MRA_FWD_SAVE_FOR_BACKWARD, "mra_qformer_backward: the forward must run with MRA_FWD_SAVE_FOR_BACKWARD");
    if (int e = device_check()) return e;
    const auto& c = h->cfg;
    const auto& W = h->w;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const int rows = io->rows, T = io->T, Nk = io->Nk, Nq = c.num_query, H = c.hidden, I = c.inter, D = c.llm_dim;
    MRA_REQUIRE(reserved == nullptr, "mra_qformer_backward: the reserved argument must be NULL");
    MRA_REQUIRE(D > 0 && W.w_proj && g->w_proj && g->b_proj, "mra_qformer_backward needs the projection (llm_dim > 0)");
    Workspace ws = carve(h, rows, T, Nk, io->flags, workspace);
    MRA_REQUIRE(workspace_bytes >= ws.total, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.total);
    BwdWorkspace bw = carve_bwd(h, rows, T, Nk, bwd_workspace);
    MRA_REQUIRE(bwd_bytes >= bw.total, "backward workspace too small: %zu < %zu bytes", bwd_bytes, bw.total);
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(bwd_workspace) & 255) == 0, "backward workspace must be 256-byte aligned");
    const int S = Nq + T;
    const int Mq = rows * Nq, Mt = rows * T, Mtot = Mq + Mt;
    const int kv_ld = h->n_cross * 2 * H;
    const int NK = rows * Nk;
    const bool skip_dead = (io->flags & MRA_FWD_SKIP_DEAD_TEXT_FFN) != 0;
    const DropoutParams drop0 = dropout_base(io);   // the forward's dropout: every mask is regenerated from (seed, site, layer)
    auto dsite = [&](int site, int layer) { return drop0.thr8 != 0 ? dropout_site(drop0, site, layer) : DropoutParams(); };
    int launches = 0;
#define MRA_TRY(expr)            \
    do {                         \
        if (int _e = (expr)) return _e; \
        ++launches;              \
    } while (0)
    // data gradient dX[M, N] = dY[M, K] . W[K, N] (+ res): W is the forward's [out = K, in = N] weight, read in place
    // (gemm.cu operand form 2); the ldw argument of the call sites (stride of the former transposed copies) is unused
    auto gemm = [&](const void* A, int64_t lda, const void* Wfwd, int64_t, const float* res, int64_t ldr, void* C, int64_t ldc,
                    int M, int N, int K, int f32) -> int {
        GemmArgs a{A, lda, Wfwd, N, nullptr, res, ldr, C, ldc, M, N, K, 0, f32};
        a.tn = 2;
        return launch_gemm_tc(a, s);
    };
    // dW[N_out, K_in] += dY^T X  (+ db += colsum dY):  dY bf16 [n, N_out] (ld ldy), X bf16 [n, K_in] (ld ldx).
    // The GEMM reads dY and X as they lie (MN-major descriptors, gemm.cu "TN"): no transposed copies.
    auto wgrad = [&](const __nv_bfloat16* dY, int64_t ldy, const __nv_bfloat16* X, int64_t ldx, int n, int N_out, int K_in,
                     float* gW, int64_t ldg, float* gB) -> int {
        if (gB != nullptr) MRA_TRY(launch_colsum(dY, ldy, n, N_out, gB, s));
        GemmArgs a{dY, ldy, X, ldx, nullptr, gW, ldg, gW, ldg, N_out, K_in, n, 0, 1};
        a.tn = 1;
        MRA_TRY(launch_gemm_tc(a, s));   // (the SIMT debug kernel has no transposed-operand form)
        return 0;
    };
    // LayerNorm backward of a post-LN block; dbias = bias gradient of the Linear feeding the LayerNorm (fused column sums)
    auto ln_bwd = [&](const float* dy, const float* pre, const float* gamma, float* dgam, float* dbet, float* dbias, size_t r0,
                      int n, const DropoutParams& dout, const DropoutParams& din) -> int {
        const size_t o = r0 * H;
        MRA_TRY(launch_ln_bwd(dy + o, pre + o, gamma, bw.g_pre32 + o, bw.g_pre16 + o, dgam, dbet, dbias, n, H, c.ln_eps, s,
                              dout.thr8 ? &dout : nullptr, din.thr8 ? &din : nullptr, static_cast<int>(r0)));
        return 0;
    };
    const DropoutParams no_drop;
    // FFN backward over rows [r0, r0+n): g_out = bw.g_x rows -> g_in accumulated into bw.g_a rows
    auto ffn_bwd = [&](const LayerBufs& B, const __nv_bfloat16* in16, size_t r0, int n, const void* w1T, const void* w2T,
                       const float* gamma, float* g_w1, float* g_b1, float* g_w2, float* g_b2, float* g_gam, float* g_bet,
                       int layer) -> int {
        const size_t o = r0 * H, oi = r0 * I;
        if (int e = ln_bwd(bw.g_x, B.pre_f, gamma, g_gam, g_bet, g_b2, r0, n, dsite(DROP_FFN_OUT, layer), no_drop)) return e;
        if (int e = wgrad(bw.g_pre16 + o, H, B.inter + oi, I, n, H, I, g_w2, I, nullptr)) return e;
        MRA_TRY(gemm(bw.g_pre16 + o, H, w2T, H, nullptr, 0, bw.g_big16, I, n, I, H, 0));          // d_inter
        if (g_b1 != nullptr) MRA_TRY(launch_gelu_bwd_colsum(B.z + oi, bw.g_big16, bw.g_big2, g_b1, n, I, s));   // dz (+ db1)
        else MRA_TRY(launch_gelu_bwd(B.z + oi, bw.g_big16, bw.g_big2, static_cast<int64_t>(n) * I, s));
        if (int e = wgrad(bw.g_big2, I, in16 + o, H, n, I, H, g_w1, H, nullptr)) return e;
        MRA_TRY(gemm(bw.g_big2, I, w1T, I, bw.g_pre32 + o, H, bw.g_a + o, H, n, H, I, 1));          // + residual path
        return 0;
    };

    // FFN_query and FFN_text backward of one layer TOGETHER (rows [0, Mq) and [Mq, Mq + Mt), Mq == Mt): the two chains are
    // independent and have the same shapes, so every GEMM of the pair is ONE grouped launch (gemm.cu takes up to four problems
    // with their own operands): 4 launches fewer per layer; the step is bound by its ~800 small launches (profiles/r02_NOTES.md)
    auto gemm2 = [&](GemmArgs a0, GemmArgs a1) -> int {
        GemmArgs pair[2] = {a0, a1};
        return launch_gemm_tc_grouped(pair, 2, s);
    };
    auto ffn_bwd_pair = [&](const LayerBufs& B, const __nv_bfloat16* in16_q, const mra_qformer_layer_weights& LW, const mra_qformer_layer_grads& G,
                            int layer) -> int {
        const size_t o = static_cast<size_t>(Mq) * H, oi = static_cast<size_t>(Mq) * I;
        const DropoutParams dout = dsite(DROP_FFN_OUT, layer);
        if (int e = ln_bwd(bw.g_x, B.pre_f, LW.ln_fq_g, G.ln_fq_g, G.ln_fq_b, G.b_fq2, 0, Mq, dout, no_drop)) return e;
        if (int e = ln_bwd(bw.g_x, B.pre_f, LW.ln_ft_g, G.ln_ft_g, G.ln_ft_b, G.b_ft2, Mq, Mt, dout, no_drop)) return e;
        {   // dW2 += d_pre^T inter  (query | text)
            GemmArgs a{bw.g_pre16, H, B.inter, I, nullptr, G.w_fq2, I, G.w_fq2, I, H, I, Mq, 0, 1};
            GemmArgs b{bw.g_pre16 + o, H, B.inter + oi, I, nullptr, G.w_ft2, I, G.w_ft2, I, H, I, Mt, 0, 1};
            a.tn = b.tn = 1;
            MRA_TRY(gemm2(a, b));
        }
        {   // d_inter = d_pre W2
            GemmArgs a{bw.g_pre16, H, LW.w_fq2, I, nullptr, nullptr, 0, bw.g_big16, I, Mq, I, H, 0, 0};
            GemmArgs b{bw.g_pre16 + o, H, LW.w_ft2, I, nullptr, nullptr, 0, bw.g_big16 + oi, I, Mt, I, H, 0, 0};
            a.tn = b.tn = 2;
            MRA_TRY(gemm2(a, b));
        }
        if (G.b_fq1 != nullptr) MRA_TRY(launch_gelu_bwd_colsum(B.z, bw.g_big16, bw.g_big2, G.b_fq1, Mq, I, s));   // dz (+ db1)
        else MRA_TRY(launch_gelu_bwd(B.z, bw.g_big16, bw.g_big2, static_cast<int64_t>(Mq) * I, s));
        if (G.b_ft1 != nullptr) MRA_TRY(launch_gelu_bwd_colsum(B.z + oi, bw.g_big16 + oi, bw.g_big2 + oi, G.b_ft1, Mt, I, s));
        else MRA_TRY(launch_gelu_bwd(B.z + oi, bw.g_big16 + oi, bw.g_big2 + oi, static_cast<int64_t>(Mt) * I, s));
        {   // dW1 += dz^T x
            GemmArgs a{bw.g_big2, I, in16_q, H, nullptr, G.w_fq1, H, G.w_fq1, H, I, H, Mq, 0, 1};
            GemmArgs b{bw.g_big2 + oi, I, B.ab + o, H, nullptr, G.w_ft1, H, G.w_ft1, H, I, H, Mt, 0, 1};
            a.tn = b.tn = 1;
            MRA_TRY(gemm2(a, b));
        }
        {   // d_x = dz W1 + the residual path
            GemmArgs a{bw.g_big2, I, LW.w_fq1, H, nullptr, bw.g_pre32, H, bw.g_a, H, Mq, H, I, 0, 1};
            GemmArgs b{bw.g_big2 + oi, I, LW.w_ft1, H, nullptr, bw.g_pre32 + o, H, bw.g_a + o, H, Mt, H, I, 0, 1};
            a.tn = b.tn = 2;
            MRA_TRY(gemm2(a, b));
        }
        return 0;
    };
    static const bool pair_ffn = [] { const char* e = getenv("MRA_BWD_PAIR_FFN"); return e == nullptr || atoi(e) != 0; }();

    // ---- llm_proj
    const __nv_bfloat16* dl = reinterpret_cast<const __nv_bfloat16*>(d_llm);
    if (int e = wgrad(dl, D, ws.layer[c.layers].xb, H, Mq, D, H, g->w_proj, H, g->b_proj)) return e;
    MRA_TRY(gemm(dl, D, W.w_proj, D, nullptr, 0, bw.g_x, H, Mq, H, D, 1));
    // projection gradients final AND its weights no longer read by this backward (the optimizer may rewrite them)
    if (h->layer_done[c.layers]) MRA_CHECK_CUDA(cudaEventRecord(h->layer_done[c.layers], s));
    if (Mt > 0) {
        MRA_CHECK_CUDA(cudaMemsetAsync(bw.g_x + static_cast<size_t>(Mq) * H, 0, static_cast<size_t>(Mt) * H * 4, s));
        ++launches;
    }

    for (int l = c.layers - 1; l >= 0; --l) {
        const auto& L = W.layer[l];
        const auto& LT = W.layer[l];   // forward weights, read in place by the data-gradient GEMMs
        const auto& G = g->layer[l];
        const LayerBufs& B = ws.layer[l];
        const bool last = l == c.layers - 1;
        const bool cross = h->cross_slot[l] >= 0;
        // ---- FFN_query (rows < Mq) and FFN_text
        const bool text_ffn = Mt > 0 && !(last && skip_dead);
        if (pair_ffn && text_ffn && Mt == Mq) {
            MRA_REQUIRE(LT.w_ft1 && LT.w_ft2 && G.w_ft1 && G.w_ft2, "layer %d: text FFN weights / grads missing", l);
            if (int e = ffn_bwd_pair(B, cross ? B.ab2 : B.ab, L, G, l)) return e;
        } else {
        if (int e = ffn_bwd(B, cross ? B.ab2 : B.ab, 0, Mq, LT.w_fq1, LT.w_fq2, L.ln_fq_g, G.w_fq1, G.b_fq1, G.w_fq2, G.b_fq2,
                            G.ln_fq_g, G.ln_fq_b, l)) return e;
        if (Mt > 0) {
            const size_t o = static_cast<size_t>(Mq) * H;
            if (last && skip_dead) {
                MRA_CHECK_CUDA(cudaMemcpyAsync(bw.g_a + o, bw.g_x + o, static_cast<size_t>(Mt) * H * 4, cudaMemcpyDeviceToDevice, s));
                ++launches;
            } else {
                MRA_REQUIRE(LT.w_ft1 && LT.w_ft2 && G.w_ft1 && G.w_ft2, "layer %d: text FFN weights / grads missing", l);
                if (int e = ffn_bwd(B, B.ab, Mq, Mt, LT.w_ft1, LT.w_ft2, L.ln_ft_g, G.w_ft1, G.b_ft1, G.w_ft2, G.b_ft2,
                                    G.ln_ft_g, G.ln_ft_b, l)) return e;
            }
        }
        }
        // ---- cross-attention block (query rows): g_a[:Mq] is the gradient w.r.t. LN_c's output
        if (cross) {
            const int slot = h->cross_slot[l];
            const __nv_bfloat16* kbase = ws.kv + static_cast<size_t>(slot) * 2 * H;
            if (int e = ln_bwd(bw.g_a, B.pre_c, L.ln_c_g, G.ln_c_g, G.ln_c_b, G.b_co, 0, Mq, dsite(DROP_CROSS_OUT, l), no_drop)) return e;
            if (int e = wgrad(bw.g_pre16, H, B.cctx, H, Mq, H, H, G.w_co, H, nullptr)) return e;
            MRA_TRY(gemm(bw.g_pre16, H, LT.w_co, H, nullptr, 0, bw.g_ctx16, H, Mq, H, H, 0));      // d_cctx
            __nv_bfloat16* gkv = bw.g_kv16 + static_cast<size_t>(slot) * 2 * H;   // this layer's columns of the stacked buffer
            AttnBwdArgs a{B.cq, H, kbase, kv_ld, kbase + H, kv_ld, bw.g_ctx16, H, bw.g_cq16, H, gkv, kv_ld,
                          gkv + H, kv_ld, io->enc_mask ? ws.enc_mask : nullptr, rows, c.heads, Nq, Nk, Nq, 1};
            a.o = B.cctx; a.ldof = H;
            a.drop = dsite(DROP_CROSS_PROBS, l);
            a.db_q = G.b_cq;
            a.db_k = g->b_ckv + static_cast<size_t>(slot) * 2 * H;
            a.db_v = a.db_k + H;
            a.n_q_tokens = Mq; a.n_k_tokens = NK;
            MRA_TRY(launch_attention_bwd(a, s));
            if (int e = wgrad(bw.g_cq16, H, B.ab, H, Mq, H, H, G.w_cq, H, nullptr)) return e;
            MRA_TRY(gemm(bw.g_cq16, H, LT.w_cq, H, bw.g_pre32, H, bw.g_a, H, Mq, H, H, 1));         // -> grad of LN_a out (query rows)
        }
        // ---- self-attention block (all rows)
        if (int e = ln_bwd(bw.g_a, B.pre_a, L.ln_a_g, G.ln_a_g, G.ln_a_b, G.b_ao, 0, Mtot, dsite(DROP_SELF_OUT, l), no_drop)) return e;
        if (int e = wgrad(bw.g_pre16, H, B.ctx, H, Mtot, H, H, G.w_ao, H, nullptr)) return e;
        MRA_TRY(gemm(bw.g_pre16, H, LT.w_ao, H, nullptr, 0, bw.g_ctx16, H, Mtot, H, H, 0));         // d_ctx
        {
            AttnBwdArgs a{B.qkv, 3 * H, B.qkv + H, 3 * H, B.qkv + 2 * H, 3 * H, bw.g_ctx16, H, bw.g_big16, 3 * H,
                          bw.g_big16 + H, 3 * H, bw.g_big16 + 2 * H, 3 * H, io->attn_mask ? ws.self_mask : nullptr,
                          rows, c.heads, S, S, Nq, 0};
            a.o = B.ctx; a.ldof = H;
            a.drop = dsite(DROP_SELF_PROBS, l);
            a.db_q = G.b_qkv; a.db_k = G.b_qkv + H; a.db_v = G.b_qkv + 2 * H;
            a.n_q_tokens = a.n_k_tokens = Mtot;
            MRA_TRY(launch_attention_bwd(a, s));
        }
        if (int e = wgrad(bw.g_big16, 3 * H, B.xb, H, Mtot, 3 * H, H, G.w_qkv, H, nullptr)) return e;
        MRA_TRY(gemm(bw.g_big16, 3 * H, LT.w_qkv, 3 * H, bw.g_pre32, H, bw.g_x, H, Mtot, H, 3 * H, 1));  // grad of the layer input
        // the gradients of layers >= l (except the stacked cross K/V weights) are final: a bucketed all-reduce may start
        if (h->layer_done[l]) MRA_CHECK_CUDA(cudaEventRecord(h->layer_done[l], s));
    }
    // ---- stacked cross-attention K/V weights of ALL cross layers: dW_ckv [ncross*2H, W] += g_kv^T . enc in ONE GEMM (the
    //      encoder tokens are read once, as in the forward, and the launch fills the machine: ncross*2H x W output tiles)
    if (h->n_cross > 0) {
        if (int e = wgrad(bw.g_kv16, kv_ld, reinterpret_cast<const __nv_bfloat16*>(io->enc), c.enc_width, NK, kv_ld, c.enc_width,
                          g->w_ckv, c.enc_width, nullptr)) return e;
    }
    // ---- embeddings
    if (int e = ln_bwd(bw.g_x, ws.pre_e, W.ln_e_g, g->ln_e_g, g->ln_e_b, nullptr, 0, Mtot, no_drop, dsite(DROP_EMB, 0))) return e;
    MRA_TRY(launch_embed_bwd(bw.g_pre32, io->input_ids, g->query_tokens, io->q_rows, g->word_emb, g->pos_emb, rows, Nq, T, H,
                             c.vocab, s));
#undef MRA_TRY
    h->last_launches = launches;
    return 0;
}

extern "C" int mra_qformer_backward_layer_events(mra_qformer_t* h, void* const* events, int32_t n) {
    MRA_REQUIRE(h != nullptr && n >= 0 && n <= MRA_MAX_LAYERS + 1 && (events != nullptr || n == 0), "mra_qformer_backward_layer_events: bad arguments");
    for (int l = 0; l <= MRA_MAX_LAYERS; ++l) h->layer_done[l] = l < n ? reinterpret_cast<cudaEvent_t>(events[l]) : nullptr;
    return 0;
}

extern "C" int mra_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                             void* stream) {
    MRA_REQUIRE(params && grads && exp_avg && exp_avg_sq, "mra_adam_step: NULL argument");
    if (int e = device_check()) return e;
    return launch_adam(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                       reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_adam_step_fused(float* params, float* grads, const void* reduced_grads_bf16, float* exp_avg, float* exp_avg_sq,
                                   void* params_bf16, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                   int32_t step, float grad_scale, int32_t zero_grads, void* stream) {
    MRA_REQUIRE(params && grads && exp_avg && exp_avg_sq, "mra_adam_step_fused: NULL argument");
    if (int e = device_check()) return e;
    return launch_adam_fused(params, grads, reduced_grads_bf16, exp_avg, exp_avg_sq, params_bf16, n, lr, beta1, beta2, eps,
                             weight_decay, step, grad_scale, zero_grads, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mra_adam_step_fused_dyn(float* params, float* grads, const void* reduced_grads_bf16, float* exp_avg, float* exp_avg_sq,
                                       void* params_bf16, int64_t n, float beta1, float beta2, float eps, float weight_decay,
                                       int32_t zero_grads, const float* hyper_dev, void* stream) {
    MRA_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper_dev, "mra_adam_step_fused_dyn: NULL argument");
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(hyper_dev) & 15) == 0, "mra_adam_step_fused_dyn: hyper_dev must be 16-byte aligned");
    if (int e = device_check()) return e;
    return launch_adam_fused(params, grads, reduced_grads_bf16, exp_avg, exp_avg_sq, params_bf16, n, 0.f, beta1, beta2, eps,
                             weight_decay, 1, 1.f, zero_grads, reinterpret_cast<cudaStream_t>(stream), hyper_dev);
}

extern "C" void mra_adam_hyper(float lr, float beta1, float beta2, int32_t step, float grad_scale, float* out4) {
    if (out4) adam_hyper(lr, beta1, beta2, step, grad_scale, out4);
}

extern "C" int mra_cast_bf16(const float* in, void* out, int64_t n, void* stream) {
    MRA_REQUIRE(in && out, "mra_cast_bf16: NULL argument");
    if (int e = device_check()) return e;
    return launch_cast_bf16(in, out, n, reinterpret_cast<cudaStream_t>(stream));
}
