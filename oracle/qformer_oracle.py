"""CPU fp32 oracle for the Q-Former + llm_proj hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module; the product (``mraudio_b200``) never does.

What it restates
----------------
The reference (globc/mrAudio) does not contain the Q-Former arithmetic; it imports it from the
un-vendored, un-pinned dependency ``salesforce/LAVIS`` (``lavis/models/blip2_models/Qformer.py`` on top of
``transformers==4.33.2``) at ``/root/reference/models/xinstructblip.py:15`` and calls it at ``:286-293`` and
``:461-468``.  This file is a self-contained pure-PyTorch restatement of that published algorithm
(BLIP-2 / InstructBLIP Q-Former: post-LN BERT, cross-attention every ``cross_attention_freq`` layers on the
query tokens only, separate query/text FFNs, erf-GELU, additive -10000 text mask) plus the reference's own
glue:

* ``init_qformer_weights``   <- ``models/xinstructblip.py:615-628`` (config + query-token init N(0, 0.02)),
                                BERT ``_init_weights`` (Linear/Embedding N(0, 0.02), bias 0, LN 1/0)
* ``qformer_bert``           <- ``Qformer.bert(...)`` call at ``models/xinstructblip.py:286-293``
* ``llm_proj``               <- ``models/xinstructblip.py:303,475`` (``nn.Linear(768, 4096)``, ctor ``:707-708``)
* ``modality_layernorm``     <- ``models/xinstructblip.py:822-828`` (fp32-upcast LayerNorm)
* ``xinstructblip_encode``   <- ``models/xinstructblip.py:280-305`` (frame fold, batch-major reorder, frame-major text
                                tiling -- including the text/visual row mismatch for bs > 1 -- slice, proj, reshape)
* ``videollama_v1_encode``   <- Video-LLaMA v1 video Q-Former (frame position embedding); NOT in the reference or its
                                pinned deps (``models/videollama.py:1-25`` wraps VideoLLaMA2) => parity unpinned.

Pinning
-------
The reference holds no tests or golden vectors for this path and LAVIS is not installable
offline, so the oracle is pinned against the HuggingFace port of the same arithmetic that IS installed here
(``transformers.models.instructblip.modeling_instructblip.InstructBlipQFormerModel`` and
``transformers.models.blip_2.modeling_blip_2.Blip2QFormerModel``): ``tests/test_oracle_qformer.py`` checks
agreement to 1e-5, and ``tests/golden/make_golden.py`` froze outputs of the HF module into
``tests/golden/qformer_*.npz``.  With respect to the reference's *own* artefacts the parity is "unpinned".

State-dict keys follow LAVIS naming (``bert.encoder.layer.{i}.attention.self.query.weight`` ...).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F


@dataclass
class QFormerOracleConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    encoder_width: int = 1408
    cross_attention_freq: int = 2
    query_length: int = 32
    vocab_size: int = 30523          # bert-base-uncased 30522 + [DEC]  (models/xinstructblip.py:622,134)
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12
    initializer_range: float = 0.02
    has_text: bool = True            # False => query-only Q-Former (no word/position embeddings used)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def has_cross(self, layer: int) -> bool:
        return layer % self.cross_attention_freq == 0


def _normal(gen, *shape, std):
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std


def init_qformer_weights(cfg: QFormerOracleConfig, seed: int = 0, llm_dim: int = 4096,
                         randomize_ln_and_bias: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights with LAVIS key names.  ``randomize_ln_and_bias`` perturbs LN gamma/beta and biases away
    from the (1, 0, 0) of BERT init so tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    H, I, W = cfg.hidden_size, cfg.intermediate_size, cfg.encoder_width
    std = cfg.initializer_range
    w: Dict[str, torch.Tensor] = {}

    def lin(prefix, out_f, in_f):
        w[prefix + ".weight"] = _normal(g, out_f, in_f, std=std)
        w[prefix + ".bias"] = (_normal(g, out_f, std=std) if randomize_ln_and_bias else torch.zeros(out_f))

    def ln(prefix, n):
        if randomize_ln_and_bias:
            w[prefix + ".weight"] = 1.0 + _normal(g, n, std=0.1)
            w[prefix + ".bias"] = _normal(g, n, std=0.05)
        else:
            w[prefix + ".weight"] = torch.ones(n)
            w[prefix + ".bias"] = torch.zeros(n)

    w["bert.embeddings.word_embeddings.weight"] = _normal(g, cfg.vocab_size, H, std=std)
    w["bert.embeddings.position_embeddings.weight"] = _normal(g, cfg.max_position_embeddings, H, std=std)
    ln("bert.embeddings.LayerNorm", H)
    for i in range(cfg.num_hidden_layers):
        p = f"bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            lin(p + f"attention.self.{n}", H, H)
        lin(p + "attention.output.dense", H, H)
        ln(p + "attention.output.LayerNorm", H)
        if cfg.has_cross(i):
            lin(p + "crossattention.self.query", H, H)
            lin(p + "crossattention.self.key", H, W)
            lin(p + "crossattention.self.value", H, W)
            lin(p + "crossattention.output.dense", H, H)
            ln(p + "crossattention.output.LayerNorm", H)
        lin(p + "intermediate.dense", I, H)
        lin(p + "output.dense", H, I)
        ln(p + "output.LayerNorm", H)
        lin(p + "intermediate_query.dense", I, H)
        lin(p + "output_query.dense", H, I)
        ln(p + "output_query.LayerNorm", H)
    # module-level tensors of XInstructBLIP (models/xinstructblip.py:624-627, 707-708, 678-704)
    w["query_tokens"] = _normal(g, 1, cfg.query_length, H, std=std)
    bound = 1.0 / math.sqrt(H)  # nn.Linear default init range
    w["llm_proj.weight"] = (torch.rand(llm_dim, H, generator=g) * 2 - 1) * bound
    w["llm_proj.bias"] = (torch.rand(llm_dim, generator=g) * 2 - 1) * bound
    ln("ln", W)
    return w


# ---------------------------------------------------------------------------------------------------------------------
# bf16 emulation: when ``emu`` is True every tensor the CUDA path stores in bf16 is rounded here at the same point, so
# the GPU result can be compared with a tight tolerance; with ``emu`` False this is the plain fp32 oracle.
# ---------------------------------------------------------------------------------------------------------------------
def _on(emu, tag: str) -> bool:
    if isinstance(emu, (set, frozenset)):
        return tag in emu
    return bool(emu) and tag in ("w", "act")


def _r(x: torch.Tensor, emu, tag: str = "act") -> torch.Tensor:
    """Round to bf16 and back when emulation is on.  ``emu`` is True (= {"w", "act"}: what the CUDA path stores in
    bf16), False, or an explicit set of tags (used to study which storage points dominate the error): "w" weights,
    "act" GEMM operands/outputs and attention outputs, "pre" pre-LN residual sums, "ln" LN outputs (residual stream)."""
    return x.to(torch.bfloat16).to(torch.float32) if _on(emu, tag) else x


def _linear(x, w, b, emu):
    return F.linear(x, _r(w, emu, "w"), b)


def _layernorm(x, g, b, eps):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


# ---------------------------------------------------------------------------------------------------------------------
# Training-mode dropout (model.train() of utils/trainer.py:110; HF port modeling_instructblip.py:530,551,608,781).  The CUDA
# path draws its masks from a counter-based Philox4x32-10 stream (mraudio_b200/csrc/dropout.cuh); this is the same generator
# in numpy with the same (seed, site, layer, element) -> byte mapping, so forward and gradients can be compared with the
# IDENTICAL mask.  Sites: 1 embeddings, 2 self-attention probs, 3 cross-attention probs, 4 self-output, 5 cross-output,
# 6 FFN output.
# ---------------------------------------------------------------------------------------------------------------------
DROP_EMB, DROP_SELF_PROBS, DROP_CROSS_PROBS, DROP_SELF_OUT, DROP_CROSS_OUT, DROP_FFN_OUT = 1, 2, 3, 4, 5, 6


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    import numpy as np
    M = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & M for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & M, np.uint64(k1) & M
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M, p1 >> np.uint64(32), p1 & M
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return c0, c1, c2, c3


def dropout_bytes(seed: int, site: int, layer: int, idx):
    """uint8 [..., 16]: the 16 random bytes of call ``idx`` of the stream (site, layer) -- dropout.cuh::dropout_bytes"""
    import numpy as np
    idx = np.asarray(idx, dtype=np.uint64)
    w = _philox4x32_10(idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full(idx.shape, site * 256 + layer, np.uint64),
                       np.full(idx.shape, 0x6d72, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty(idx.shape + (16,), dtype=np.uint8)
    for k in range(16):
        out[..., k] = ((w[k // 4] >> np.uint64(8 * (k % 4))) & np.uint64(0xFF)).astype(np.uint8)
    return out


def dropout_params(p: float):
    """(thr8, scale): an element is kept iff its byte >= thr8 = round(256 p); kept values are scaled by 256 / (256 - thr8)"""
    thr = min(255, max(1, int(p * 256.0 + 0.5)))
    return thr, 256.0 / (256 - thr)


def hidden_dropout_mult(p: float, seed: int, site: int, layer: int, m_index: torch.Tensor, H: int) -> torch.Tensor:
    """multiplier [..., H] for tokens whose split-layout row indices are ``m_index`` [...] (queries of all rows first, then
    text): element (m, n) -> call m * (H / 8) + n / 8, byte n % 8"""
    import numpy as np
    thr, scale = dropout_params(p)
    m = m_index.numpy().astype(np.uint64)
    idx = m[..., None] * np.uint64(H // 8) + np.arange(H // 8, dtype=np.uint64)
    b = dropout_bytes(seed, site, layer, idx)[..., :8]                 # [..., H/8, 8]
    keep = (b >= thr).reshape(m.shape + (H,))
    return torch.from_numpy(keep.astype(np.float32) * np.float32(scale))


def attention_dropout_mult(p: float, seed: int, site: int, layer: int, rows: int, heads: int, Sq: int, Sk: int) -> torch.Tensor:
    """multiplier [rows, heads, Sq, Sk]: element (r, h, i, j) -> call (((r heads + h) Sq + i) ceil(Sk / 64) + j / 64) 4 + (j % 8) / 2,
    byte ((j % 64) / 8) 2 + j % 2"""
    import numpy as np
    thr, scale = dropout_params(p)
    nch = (Sk + 63) // 64
    r, h, i, j = np.meshgrid(np.arange(rows), np.arange(heads), np.arange(Sq), np.arange(Sk), indexing="ij")
    idx = (((r * heads + h) * Sq + i) * nch + j // 64) * 4 + (j % 8) // 2
    byte = ((j % 64) // 8) * 2 + j % 2
    b = dropout_bytes(seed, site, layer, idx.astype(np.uint64))
    sel = np.take_along_axis(b, byte[..., None].astype(np.int64), axis=-1)[..., 0]
    return torch.from_numpy((sel >= thr).astype(np.float32) * np.float32(scale))


def _attention(q, k, v, add_mask, nheads, emu, drop_mult=None):
    """q:[R,Sq,H] k,v:[R,Sk,H] add_mask:[R,Sk] or None -> [R,Sq,H].  scores/8 + mask -> softmax -> (dropout) -> PV."""
    R, Sq, H = q.shape
    Sk = k.shape[1]
    d = H // nheads
    qh = q.view(R, Sq, nheads, d).permute(0, 2, 1, 3)
    kh = k.view(R, Sk, nheads, d).permute(0, 2, 1, 3)
    vh = v.view(R, Sk, nheads, d).permute(0, 2, 1, 3)
    s = torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(d)
    if add_mask is not None:
        s = s + add_mask[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    if _on(emu, "act"):
        # the kernel feeds un-normalised bf16 probabilities to the PV tensor-core product and divides by the fp32
        # row sum afterwards
        m = s.max(dim=-1, keepdim=True).values
        e = torch.exp(s - m)
        ed = e if drop_mult is None else e * drop_mult
        o = torch.matmul(_r(ed, True), vh) / e.sum(dim=-1, keepdim=True)
    else:
        o = torch.matmul(p if drop_mult is None else p * drop_mult, vh)
    return o.permute(0, 2, 1, 3).reshape(R, Sq, H)


def qformer_bert(w: Dict[str, torch.Tensor], cfg: QFormerOracleConfig,
                 input_ids: Optional[torch.Tensor], attention_mask: Optional[torch.Tensor],
                 query_embeds: torch.Tensor, encoder_hidden_states: torch.Tensor,
                 encoder_attention_mask: Optional[torch.Tensor] = None,
                 emulate_bf16: bool = False, skip_dead_text_ffn: bool = False, dropout=None) -> torch.Tensor:
    """``Qformer.bert(input_ids, attention_mask=, query_embeds=, encoder_hidden_states=, encoder_attention_mask=,
    return_dict=True).last_hidden_state`` -> [rows, Nq+T, H]   (call site models/xinstructblip.py:286-293)."""
    emu = emulate_bf16
    eps = cfg.layer_norm_eps
    nh = cfg.num_attention_heads
    Nq = query_embeds.shape[1]
    rows = encoder_hidden_states.shape[0]
    query_embeds = query_embeds.expand(rows, -1, -1).to(torch.float32)
    enc = encoder_hidden_states.to(torch.float32)
    H = cfg.hidden_size
    Tn = input_ids.shape[1] if (input_ids is not None and cfg.has_text) else 0
    # ``dropout = (p, seed)``: training mode with the CUDA path's counter-based masks (see above); None = eval
    r_idx = torch.arange(rows)[:, None]
    m_query = r_idx * Nq + torch.arange(Nq)[None, :]                                   # split-layout row of query token (r, i)
    m_text = rows * Nq + r_idx * Tn + torch.arange(Tn)[None, :]                        # ... of text token (r, j)
    m_all = torch.cat([m_query, m_text], dim=1)

    def hdrop(site, layer, m_index):
        return None if dropout is None else hidden_dropout_mult(dropout[0], dropout[1], site, layer, m_index, H)

    def adrop(site, layer, Sq, Sk):
        return None if dropout is None else attention_dropout_mult(dropout[0], dropout[1], site, layer, rows, nh, Sq, Sk)

    def dropped(y, mult):
        return y if mult is None else y * mult

    # --- embeddings: LN(cat(query_embeds, word_emb[ids] + pos_emb[0:T]))
    if input_ids is not None and cfg.has_text:
        T = input_ids.shape[1]
        we = _r(w["bert.embeddings.word_embeddings.weight"], emu, "w")[input_ids]
        pe = _r(w["bert.embeddings.position_embeddings.weight"], emu, "w")[:T]
        x = torch.cat([query_embeds, we + pe[None]], dim=1)
    else:
        T = 0
        x = query_embeds
    # residual stream x / a / aq stays fp32 (the CUDA path keeps an fp32 copy next to the bf16 GEMM operand copy);
    # tag "ln" is only for the error study in DESIGN.md
    x = _layernorm(x, w["bert.embeddings.LayerNorm.weight"], w["bert.embeddings.LayerNorm.bias"], eps)
    x = _r(dropped(x, hdrop(DROP_EMB, 0, m_all)), emu, "ln")

    # --- masks: (1 - m) * -10000 on keys (LAVIS get_extended_attention_mask); encoder mask inverted the same way
    if attention_mask is None:
        self_mask = None
    else:
        self_mask = (1.0 - attention_mask.to(torch.float32)) * -10000.0
    if encoder_attention_mask is None:
        enc_mask = None
    else:
        enc_mask = (1.0 - encoder_attention_mask.to(torch.float32)) * -10000.0

    def lin(inp, name):
        return _linear(_r(inp, emu), w[name + ".weight"], w[name + ".bias"], emu)

    def ln(pre, name):
        return _r(_layernorm(_r(pre, emu, "pre"), w[name + ".weight"], w[name + ".bias"], eps), emu, "ln")

    for i in range(cfg.num_hidden_layers):
        p = f"bert.encoder.layer.{i}."
        last = i == cfg.num_hidden_layers - 1
        # self-attention over queries || text
        q = _r(lin(x, p + "attention.self.query"), emu)
        k = _r(lin(x, p + "attention.self.key"), emu)
        v = _r(lin(x, p + "attention.self.value"), emu)
        ctx = _attention(q, k, v, self_mask, nh, emu, adrop(DROP_SELF_PROBS, i, Nq + T, Nq + T))
        a = ln(dropped(lin(ctx, p + "attention.output.dense"), hdrop(DROP_SELF_OUT, i, m_all)) + x, p + "attention.output.LayerNorm")

        aq = a[:, :Nq]
        if cfg.has_cross(i):
            cq = _r(lin(aq, p + "crossattention.self.query"), emu)
            ck = _r(lin(enc, p + "crossattention.self.key"), emu)
            cv = _r(lin(enc, p + "crossattention.self.value"), emu)
            cctx = _attention(cq, ck, cv, enc_mask, nh, emu, adrop(DROP_CROSS_PROBS, i, Nq, enc.shape[1]))
            aq = ln(dropped(lin(cctx, p + "crossattention.output.dense"), hdrop(DROP_CROSS_OUT, i, m_query)) + aq,
                    p + "crossattention.output.LayerNorm")

        # FFN_query
        h = F.gelu(lin(aq, p + "intermediate_query.dense"))
        yq = ln(dropped(lin(h, p + "output_query.dense"), hdrop(DROP_FFN_OUT, i, m_query)) + aq, p + "output_query.LayerNorm")
        if T > 0:
            at = a[:, Nq:]
            if last and skip_dead_text_ffn:
                yt = at  # layer-11 text FFN output is never read by the hot path (models/xinstructblip.py:303)
            else:
                h = F.gelu(lin(at, p + "intermediate.dense"))
                yt = ln(dropped(lin(h, p + "output.dense"), hdrop(DROP_FFN_OUT, i, m_text)) + at, p + "output.LayerNorm")
            x = torch.cat([yq, yt], dim=1)
        else:
            x = yq
    return x


def llm_proj(w: Dict[str, torch.Tensor], x: torch.Tensor, emulate_bf16: bool = False) -> torch.Tensor:
    """``{modality}_llm_proj(last_hidden_state[:, :32, :])``  (models/xinstructblip.py:303)."""
    return _linear(_r(x, emulate_bf16), w["llm_proj.weight"], w["llm_proj.bias"], emulate_bf16)


def modality_layernorm(w: Dict[str, torch.Tensor], x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``LayerNorm.forward``: fp32 upcast, nn.LayerNorm default eps 1e-5, cast back (models/xinstructblip.py:822-828)."""
    orig = x.dtype
    return _layernorm(x.to(torch.float32), w["ln.weight"], w["ln.bias"], eps).to(orig)


def xinstructblip_encode(w: Dict[str, torch.Tensor], cfg: QFormerOracleConfig,
                         frame_embeds: torch.Tensor, input_ids: torch.Tensor, text_mask: torch.Tensor,
                         emulate_bf16: bool = False, match_reference_text_tiling: bool = True) -> torch.Tensor:
    """One modality of ``XInstructBLIP.generate`` between the encoders and the LLM (models/xinstructblip.py:280-305).

    frame_embeds: [bs, F, Nk, W]  -- the per-frame ``ln(encoder(frame))`` outputs (already LayerNorm'ed), i.e. the
                  reference's ``embeds[modality][f][b]`` stacked.
    input_ids/text_mask: [bs, T] BERT ids / attention mask of the prompt.
    returns inputs_llm [bs, F*Nq, llm_dim].
    """
    bs, Fr, Nk, W = frame_embeds.shape
    Nq = cfg.query_length
    # frames folded batch-major: row k = b*F + f  (``torch.cat(embeds)[indices]``, :283-284)
    enc = frame_embeds.reshape(bs * Fr, Nk, W)
    if match_reference_text_tiling:
        # ``input_ids.repeat(num, 1)`` tiles frame-major: row k gets the text of sample k % bs (:287-289)
        ids = input_ids.repeat(Fr, 1)
        tmask = text_mask.repeat(Fr, 1)
    else:
        ids = input_ids.repeat_interleave(Fr, dim=0)
        tmask = text_mask.repeat_interleave(Fr, dim=0)
    q_atts = torch.ones(bs * Fr, Nq, dtype=tmask.dtype)
    atts = torch.cat([q_atts, tmask], dim=1)
    qe = w["query_tokens"].expand(bs * Fr, -1, -1)
    enc_atts = torch.ones(bs * Fr, Nk, dtype=torch.long)
    hid = qformer_bert(w, cfg, ids, atts, qe, enc, enc_atts, emulate_bf16=emulate_bf16)
    out = llm_proj(w, hid[:, :Nq, :], emulate_bf16)
    return out.reshape(bs, Fr, Nq, -1).reshape(bs, Fr * Nq, -1)


def videollama_v1_encode(w: Dict[str, torch.Tensor], cfg: QFormerOracleConfig, frame_pos_emb: torch.Tensor,
                         frame_tokens: torch.Tensor, emulate_bf16: bool = False) -> torch.Tensor:
    """Video-LLaMA-v1-style video Q-Former: per-frame tokens [B, F, n, Wd] + frame position embedding [F, Wd]
    (broadcast over the n tokens) -> keys [B, F*n, Wd] -> query-only Q-Former (cfg.has_text False) -> llm_proj.
    Not present in the reference or its pinned deps: parity unpinned."""
    B, Fr, n, Wd = frame_tokens.shape
    x = frame_tokens.to(torch.float32) + frame_pos_emb[:Fr].to(torch.float32)[None, :, None, :]
    enc = x.reshape(B, Fr * n, Wd)
    qe = w["query_tokens"].expand(B, -1, -1)
    hid = qformer_bert(w, cfg, None, None, qe, enc, None, emulate_bf16=emulate_bf16)
    return llm_proj(w, hid, emulate_bf16)


def algorithmic_flops_per_row(cfg: QFormerOracleConfig, T: int, Nk: int, llm_dim: int = 4096) -> float:
    """2*M*N*K per GEMM including QK^T and PV; softmax/LN/GELU/bias not counted."""
    H, I, W, Nq = cfg.hidden_size, cfg.intermediate_size, cfg.encoder_width, cfg.query_length
    S = Nq + T
    L = cfg.num_hidden_layers
    Lc = sum(1 for i in range(L) if cfg.has_cross(i))
    f = L * (3 * 2 * S * H * H + 2 * S * H * H + 4 * S * S * H)
    f += Lc * (2 * 2 * Nq * H * H + 2 * 2 * Nk * W * H + 4 * Nq * Nk * H)
    f += L * (4 * Nq * H * I) + L * (4 * T * H * I)
    f += 2 * Nq * H * llm_dim
    return float(f)
