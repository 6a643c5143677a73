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

struct Workspace {
    float* x32; __nv_bfloat16* xb;     // layer input  (residual stream + bf16 operand copy)
    float* a32; __nv_bfloat16* ab;     // attention output
    float* pre;                        // pre-LayerNorm sums (fp32)
    __nv_bfloat16* qkv;                // [Mtot, 3H]
    __nv_bfloat16* ctx;                // [Mtot, H]
    __nv_bfloat16* inter;              // [Mtot, I]
    __nv_bfloat16* cq;                 // [Mq, H]
    __nv_bfloat16* kv;                 // [rows*Nk, ncross*2H]
    float* self_mask;                  // [rows, S]
    float* enc_mask;                   // [rows, Nk]
    size_t total;
};

Workspace carve(const mra_qformer* h, int rows, int T, int Nk, void* base) {
    const auto& c = h->cfg;
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
    w.xb = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    w.a32 = reinterpret_cast<float*>(take(Mtot * H * 4));
    w.ab = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    w.pre = reinterpret_cast<float*>(take(Mtot * H * 4));
    w.qkv = reinterpret_cast<__nv_bfloat16*>(take(Mtot * 3 * H * 2));
    w.ctx = reinterpret_cast<__nv_bfloat16*>(take(Mtot * H * 2));
    w.inter = reinterpret_cast<__nv_bfloat16*>(take(Mtot * I * 2));
    w.cq = reinterpret_cast<__nv_bfloat16*>(take(Mq * H * 2));
    w.kv = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(rows) * Nk * h->n_cross * 2 * H * 2));
    w.self_mask = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * (c.num_query + T) * 4));
    w.enc_mask = reinterpret_cast<float*>(take(static_cast<size_t>(rows) * Nk * 4));
    w.total = off;
    return w;
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

extern "C" size_t mra_qformer_workspace_bytes(const mra_qformer_t* h, int32_t rows, int32_t T, int32_t Nk, uint32_t) {
    if (!h || rows <= 0 || T < 0 || Nk <= 0) return 0;
    return carve(h, rows, T, Nk, nullptr).total;
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

extern "C" int mra_qformer_forward(mra_qformer_t* h, const mra_qformer_io* io, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
    MRA_REQUIRE(h && io && workspace, "mra_qformer_forward: NULL argument");
    MRA_REQUIRE(h->has_weights, "mra_qformer_forward: weights not set");
    if (int e = device_check()) return e;
    const auto& c = h->cfg;
    const auto& W = h->w;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    const int rows = io->rows, T = io->T, Nk = io->Nk, Nq = c.num_query, H = c.hidden, I = c.inter;
    MRA_REQUIRE(rows > 0 && Nk > 0 && T >= 0, "bad shape rows=%d T=%d Nk=%d", rows, T, Nk);
    MRA_REQUIRE(T <= c.max_pos, "T=%d exceeds max_position_embeddings=%d", T, c.max_pos);
    MRA_REQUIRE(io->enc && io->query_embeds, "enc / query_embeds must not be NULL");
    MRA_REQUIRE(io->q_rows == 1 || io->q_rows == rows, "query_embeds rows must be 1 or %d, got %d", rows, io->q_rows);
    MRA_REQUIRE(T == 0 || (io->input_ids && W.word_emb && W.pos_emb), "text tokens need input_ids and embedding tables");
    MRA_REQUIRE(T == 0 || W.layer[0].w_ft1, "text tokens need the text FFN weights");
    MRA_REQUIRE(!io->llm_out || (W.w_proj && W.b_proj && c.llm_dim > 0), "llm_out requested but no projection weights");
    Workspace ws = carve(h, rows, T, Nk, workspace);
    MRA_REQUIRE(workspace_bytes >= ws.total, "workspace too small: %zu < %zu bytes", workspace_bytes, ws.total);
    MRA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");

    const int S = Nq + T;
    const int Mq = rows * Nq, Mt = rows * T, Mtot = Mq + Mt;
    const int kv_ld = h->n_cross * 2 * H;
    int launches = 0;
    // profiling spans: MRA_PROFILE_DOMINANT brackets only GEMM launches, MRA_PROFILE_ALL every launch
    int cur_cat = -1;
    cudaEvent_t cur_ev = nullptr;
    auto span_begin = [&](int cat) {
        cur_cat = -1;
        if (h->profile_mode == MRA_PROFILE_OFF) return;
        if (h->profile_mode == MRA_PROFILE_DOMINANT && cat > MRA_CAT_GEMM) return;
        cur_cat = cat;
        cur_ev = h->get_event();
        cudaEventRecord(cur_ev, s);
    };
    auto span_end = [&]() {
        if (cur_cat < 0) return;
        cudaEvent_t b = h->get_event();
        cudaEventRecord(b, s);
        h->spans.push_back({cur_cat, cur_ev, b});
        cur_cat = -1;
    };
    int gemm_cat = MRA_CAT_GEMM;
    auto gemm = [&](const void* A, int64_t lda, const void* Wt, int64_t ldw, const float* bias, const float* res,
                    int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int gelu, int f32) -> int {
        GemmArgs a{A, lda, Wt, ldw, bias, res, ldr, C, ldc, M, N, K, gelu, f32};
        ++launches;
        span_begin(gemm_cat);
        int e = h->gemm_impl == MRA_GEMM_IMPL_SIMT_DEBUG ? launch_gemm_simt(a, s) : launch_gemm_tc(a, s);
        span_end();
        return e;
    };
#define MRA_TRY(expr)            \
    do {                         \
        if (int _e = (expr)) return _e; \
    } while (0)

    // ---- embeddings + masks
    span_begin(MRA_CAT_OTHER);
    MRA_TRY(launch_embed_layernorm(io->query_embeds, io->q_rows, io->input_ids, W.word_emb, W.pos_emb, W.ln_e_g, W.ln_e_b,
                                   ws.x32, ws.xb, rows, Nq, T, H, c.vocab, c.ln_eps, s));
    span_end();
    ++launches;
    const float* self_mask = nullptr;
    if (io->attn_mask) {
        MRA_TRY(launch_build_enc_mask(io->attn_mask, ws.self_mask, rows, Nq + T, s));
        ++launches;
        self_mask = ws.self_mask;
    }
    const float* enc_mask = nullptr;
    if (io->enc_mask) {
        MRA_TRY(launch_build_enc_mask(io->enc_mask, ws.enc_mask, rows, Nk, s));
        ++launches;
        enc_mask = ws.enc_mask;
    }
    // ---- cross-attention keys / values of ALL cross layers in one GEMM: the encoder tokens are read once
    gemm_cat = MRA_CAT_GEMM_CROSS_KV;
    MRA_TRY(gemm(io->enc, c.enc_width, W.w_ckv, c.enc_width, W.b_ckv, nullptr, 0, ws.kv, kv_ld, rows * Nk, kv_ld,
                 c.enc_width, 0, 0));
    gemm_cat = MRA_CAT_GEMM;

    for (int l = 0; l < c.layers; ++l) {
        const auto& L = W.layer[l];
        const bool last = l == c.layers - 1;
        // self-attention over queries || text
        MRA_TRY(gemm(ws.xb, H, L.w_qkv, H, L.b_qkv, nullptr, 0, ws.qkv, 3 * H, Mtot, 3 * H, H, 0, 0));
        {
            AttnArgs a{ws.qkv, 3 * H, ws.qkv + H, 3 * H, ws.qkv + 2 * H, 3 * H, ws.ctx, H, self_mask, rows, c.heads, S, S, Nq, 0};
            span_begin(MRA_CAT_ATTENTION);
            MRA_TRY(launch_attention(a, s));
            span_end();
            ++launches;
        }
        MRA_TRY(gemm(ws.ctx, H, L.w_ao, H, L.b_ao, ws.x32, H, ws.pre, H, Mtot, H, H, 0, 1));
        span_begin(MRA_CAT_LAYERNORM);
        MRA_TRY(launch_layernorm(ws.pre, L.ln_a_g, L.ln_a_b, ws.a32, ws.ab, Mtot, H, c.ln_eps, s));
        span_end();
        ++launches;
        // cross-attention of the query tokens onto this row's encoder tokens
        if (h->cross_slot[l] >= 0) {
            const __nv_bfloat16* kbase = ws.kv + static_cast<size_t>(h->cross_slot[l]) * 2 * H;
            MRA_TRY(gemm(ws.ab, H, L.w_cq, H, L.b_cq, nullptr, 0, ws.cq, H, Mq, H, H, 0, 0));
            AttnArgs a{ws.cq, H, kbase, kv_ld, kbase + H, kv_ld, ws.ctx, H, enc_mask, rows, c.heads, Nq, Nk, Nq, 1};
            span_begin(MRA_CAT_ATTENTION);
            MRA_TRY(launch_attention(a, s));
            span_end();
            ++launches;
            MRA_TRY(gemm(ws.ctx, H, L.w_co, H, L.b_co, ws.a32, H, ws.pre, H, Mq, H, H, 0, 1));
            span_begin(MRA_CAT_LAYERNORM);
            MRA_TRY(launch_layernorm(ws.pre, L.ln_c_g, L.ln_c_b, ws.a32, ws.ab, Mq, H, c.ln_eps, s));
            span_end();
            ++launches;
        }
        // FFN_query on the query rows
        MRA_TRY(gemm(ws.ab, H, L.w_fq1, H, L.b_fq1, nullptr, 0, ws.inter, I, Mq, I, H, 1, 0));
        MRA_TRY(gemm(ws.inter, I, L.w_fq2, I, L.b_fq2, ws.a32, H, ws.pre, H, Mq, H, I, 0, 1));
        span_begin(MRA_CAT_LAYERNORM);
        MRA_TRY(launch_layernorm(ws.pre, L.ln_fq_g, L.ln_fq_b, ws.x32, ws.xb, Mq, H, c.ln_eps, s));
        span_end();
        ++launches;
        // FFN_text on the text rows
        if (T > 0) {
            const size_t o = static_cast<size_t>(Mq) * H;
            if (last && (io->flags & MRA_FWD_SKIP_DEAD_TEXT_FFN)) {
                // never read by llm_proj; keep last_hidden well-defined by passing the attention output through
                if (io->last_hidden) {
                    MRA_CHECK_CUDA(cudaMemcpyAsync(ws.x32 + o, ws.a32 + o, static_cast<size_t>(Mt) * H * 4,
                                                   cudaMemcpyDeviceToDevice, s));
                    ++launches;
                }
            } else {
                MRA_REQUIRE(L.w_ft1 && L.b_ft1 && L.w_ft2 && L.b_ft2 && L.ln_ft_g && L.ln_ft_b, "layer %d: text FFN weights missing", l);
                const size_t oi = static_cast<size_t>(Mq) * I;
                MRA_TRY(gemm(ws.ab + o, H, L.w_ft1, H, L.b_ft1, nullptr, 0, ws.inter + oi, I, Mt, I, H, 1, 0));
                MRA_TRY(gemm(ws.inter + oi, I, L.w_ft2, I, L.b_ft2, ws.a32 + o, H, ws.pre + o, H, Mt, H, I, 0, 1));
                span_begin(MRA_CAT_LAYERNORM);
                MRA_TRY(launch_layernorm(ws.pre + o, L.ln_ft_g, L.ln_ft_b, ws.x32 + o, ws.xb + o, Mt, H, c.ln_eps, s));
                span_end();
                ++launches;
            }
        }
    }
    // ---- outputs
    if (io->last_hidden) {
        span_begin(MRA_CAT_OTHER);
        MRA_TRY(launch_gather_last_hidden(ws.x32, io->last_hidden, rows, Nq, T, H, s));
        span_end();
        ++launches;
    }
    if (io->llm_out) {
        MRA_TRY(gemm(ws.xb, H, W.w_proj, H, W.b_proj, nullptr, 0, io->llm_out, c.llm_dim, Mq, c.llm_dim, H, 0, 0));
    }
#undef MRA_TRY
    h->last_launches = launches;
    return 0;
}
