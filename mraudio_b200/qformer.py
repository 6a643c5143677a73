"""Host-side mirror of the LAVIS Q-Former interface that ``models/xinstructblip.py`` uses, backed by the CUDA library.

Reference surface kept (file:line under /root/reference):

* ``BertConfig`` fields set at ``models/xinstructblip.py:616-622`` (``encoder_width``, ``add_cross_attention``,
  ``cross_attention_freq``, ``query_length``, ``vocab_size``) and read at ``:161,172`` (``config.hidden_size``);
* ``BertLMHeadModel(config)`` with ``.bert``, a settable ``.cls`` (``:135``) and ``resize_token_embeddings`` (``:134``);
* ``Qformer.bert(input_ids, attention_mask=, query_embeds=, encoder_hidden_states=, encoder_attention_mask=,
  return_dict=True).last_hidden_state`` (``:286-293``, ``:461-468``);
* LAVIS state-dict key names (``bert.encoder.layer.{i}.attention.self.query.weight`` ...), so checkpoints load with
  ``load_state_dict(..., strict=False)`` exactly like ``:652`` and ``named_parameters()`` feeds the freezing loop at
  ``:196-204`` and the trainer's checkpointing (``utils/trainer.py:188-196``).

The ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Embedding`` sub-modules are parameter containers only: their ``forward`` is
never called.  ``BertModelB200.forward`` packs the weights to bf16 once per parameter version and enqueues the whole
layer stack through ``mra_qformer_forward`` on the current CUDA stream.  There is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import check, current_stream, lib


class BertConfig:
    """The subset of ``transformers.BertConfig`` (bert-base-uncased) the Q-Former reads, plus the LAVIS extensions."""

    def __init__(self, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 vocab_size=30522, max_position_embeddings=512, layer_norm_eps=1e-12, initializer_range=0.02,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, pad_token_id=0,
                 encoder_width=1408, add_cross_attention=True, cross_attention_freq=2, query_length=32):
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.layer_norm_eps = layer_norm_eps
        self.initializer_range = initializer_range
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.pad_token_id = pad_token_id
        self.encoder_width = encoder_width
        self.add_cross_attention = add_cross_attention
        self.cross_attention_freq = cross_attention_freq
        self.query_length = query_length

    @classmethod
    def from_pretrained(cls, name="bert-base-uncased", **kw):
        # models/xinstructblip.py:616 -- only the architecture constants of bert-base-uncased are needed (no files)
        assert name == "bert-base-uncased", "only the bert-base-uncased architecture is built in"
        return cls(**kw)


@dataclass
class QFormerOutput:
    """``BaseModelOutputWithPoolingAndCrossAttentions`` stand-in: attribute and ``[0]`` access to last_hidden_state."""
    last_hidden_state: Optional[torch.Tensor] = None
    llm_inputs: Optional[torch.Tensor] = None  # fused llm_proj output when a projection was passed

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class _SelfAttn(nn.Module):
    def __init__(self, hidden, kv_width):
        super().__init__()
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_width, hidden)
        self.value = nn.Linear(kv_width, hidden)


class _SelfOutput(nn.Module):
    def __init__(self, in_f, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(in_f, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)


class _Attention(nn.Module):
    def __init__(self, hidden, kv_width, eps):
        super().__init__()
        self.self = _SelfAttn(hidden, kv_width)
        self.output = _SelfOutput(hidden, hidden, eps)


class _Intermediate(nn.Module):
    def __init__(self, hidden, inter):
        super().__init__()
        self.dense = nn.Linear(hidden, inter)


class _Layer(nn.Module):
    def __init__(self, cfg: BertConfig, idx: int):
        super().__init__()
        H, I, eps = cfg.hidden_size, cfg.intermediate_size, cfg.layer_norm_eps
        self.attention = _Attention(H, H, eps)
        self.has_cross_attention = cfg.add_cross_attention and idx % cfg.cross_attention_freq == 0
        if self.has_cross_attention:
            self.crossattention = _Attention(H, cfg.encoder_width, eps)
        self.intermediate = _Intermediate(H, I)
        self.output = _SelfOutput(I, H, eps)
        self.intermediate_query = _Intermediate(H, I)
        self.output_query = _SelfOutput(I, H, eps)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(cfg, i) for i in range(cfg.num_hidden_layers)])


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(cfg.max_position_embeddings).expand((1, -1)), persistent=False)


class LLMProjB200(nn.Linear):
    """``{modality}_llm_proj = nn.Linear(768, 4096)`` (models/xinstructblip.py:707-708); state-dict keys weight/bias.
    ``forward`` runs the tcgen05 GEMM; it is also handed to ``BertModelB200.forward(llm_proj=...)`` for the fused call."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import ops
        if not x.is_cuda:
            raise _lib.MraError("LLMProjB200 needs CUDA tensors (no CPU fallback)")
        w, b = _packed_proj(self)
        x2 = x.reshape(-1, x.shape[-1]).to(torch.bfloat16).contiguous()
        y = ops.linear(x2, w, b)
        return y.reshape(*x.shape[:-1], w.shape[0])


def _param_version(mod: nn.Module) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in mod.parameters())


def _packed_proj(proj: nn.Linear):
    tr = getattr(proj, "_mra_trainable", None)
    if tr is not None:   # under fine-tuning: the trainer's live bf16 operand copy / fp32 master bias (never stale)
        return tr.proj_operands()
    ver = _param_version(proj)
    cache = getattr(proj, "_mra_pack", None)
    if cache is None or cache[0] != ver:
        cache = (ver, proj.weight.detach().to(torch.bfloat16).contiguous(), proj.bias.detach().float().contiguous())
        proj._mra_pack = cache
    return cache[1], cache[2]


class BertModelB200(nn.Module):
    """``Qformer.bert``: embeddings + encoder, executed by libmraudio_b200."""

    def __init__(self, cfg: BertConfig):
        super().__init__()
        self.config = cfg
        self.embeddings = _Embeddings(cfg)
        self.encoder = _Encoder(cfg)
        self._handles = {}    # llm_dim -> C handle
        self._pack = None     # (version key, dict of packed tensors, QFormerWeights)
        self._pack_applied = set()
        self._trainable = None   # TrainableQFormer once the module is being fine-tuned (training.py)
        self._workspace = None
        self.last_launches = 0
        self._profile_mode = _lib.PROFILE_OFF
        self._init_weights()

    # BERT _init_weights: Linear / Embedding N(0, initializer_range), bias 0, LayerNorm (1, 0)
    def _init_weights(self):
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=std)
                m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(mean=0.0, std=std)
            elif isinstance(m, nn.LayerNorm):
                m.weight.data.fill_(1.0)
                m.bias.data.zero_()

    def __del__(self):
        try:
            for h in self._handles.values():
                lib.mra_qformer_destroy(h)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------ weight packing
    def _new_handle(self, llm_dim: int):
        """A fresh C handle for this configuration (the caller owns it)."""
        c = self.config
        cc = _lib.QFormerConfig(hidden=c.hidden_size, layers=c.num_hidden_layers, heads=c.num_attention_heads,
                                inter=c.intermediate_size, enc_width=c.encoder_width, cross_freq=c.cross_attention_freq,
                                num_query=c.query_length, llm_dim=llm_dim, vocab=c.vocab_size,
                                max_pos=c.max_position_embeddings, ln_eps=c.layer_norm_eps)
        hp = C.c_void_p()
        check(lib.mra_qformer_create(C.byref(cc), C.byref(hp)))
        return hp

    def _handle(self, llm_dim: int):
        """The inference handle for this projection width (the trainer has its OWN handle: ``set_weights`` on a shared
        one would silently re-point training at an inference snapshot)."""
        h = self._handles.get(llm_dim)
        if h is None:
            h = self._new_handle(llm_dim)
            self._handles[llm_dim] = h
            check(lib.mra_qformer_profile_mode(h, self._profile_mode))
        return h

    # ------------------------------------------------------------------------------------------- device-side timing
    def set_profile_mode(self, mode: int) -> None:
        """CUDA-event timing of the forward's launches (``_lib.PROFILE_OFF / PROFILE_DOMINANT / PROFILE_ALL``)."""
        self._profile_mode = mode
        for h in self._handles.values():
            check(lib.mra_qformer_profile_mode(h, mode))

    def read_profile(self):
        """(ms per category, launches per category) accumulated since the last read; synchronises on the events."""
        ms_tot, n_tot = [0.0] * len(_lib.PROFILE_CATS), [0] * len(_lib.PROFILE_CATS)
        for h in self._handles.values():
            ms = (C.c_double * len(_lib.PROFILE_CATS))()
            n = (C.c_int64 * len(_lib.PROFILE_CATS))()
            check(lib.mra_qformer_profile_read(h, ms, n))
            ms_tot = [a + b for a, b in zip(ms_tot, ms)]
            n_tot = [a + int(b) for a, b in zip(n_tot, n)]
        return ms_tot, n_tot

    def _packed(self, proj: Optional[nn.Linear]):
        tr = self._trainable
        if tr is not None and (proj is None or proj is tr.llm_proj):
            # Under fine-tuning the optimizer updates the flat master / bf16 operand buffers through raw pointers, which
            # neither moves nor re-versions the parameters: a (data_ptr, _version) cache would serve stale weights to every
            # validation pass.  Inference therefore reads the trainer's live buffers (refreshed by every adam_step).
            ver = ("trainable", id(tr))
            if self._pack is None or self._pack[0] != ver:
                self._pack = (ver, [tr], tr.weight_struct())
                self._pack_applied = set()
            return self._pack[2]
        ver = (_param_version(self), _param_version(proj) if proj is not None else None, tr.version if tr is not None else 0)
        if self._pack is not None and self._pack[0] == ver:
            return self._pack[2]
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        keep = []   # packed tensors must outlive the handle's use of their device pointers
        W = _lib.QFormerWeights()

        def put(struct, field, tensor):
            keep.append(tensor)
            setattr(struct, field, tensor.data_ptr())

        e = self.embeddings
        put(W, "word_emb", bf(e.word_embeddings.weight))
        put(W, "pos_emb", bf(e.position_embeddings.weight))
        put(W, "ln_e_g", f32(e.LayerNorm.weight))
        put(W, "ln_e_b", f32(e.LayerNorm.bias))
        ckv_w, ckv_b = [], []
        for i, L in enumerate(self.encoder.layer):
            S = W.layer[i]
            a = L.attention
            put(S, "w_qkv", bf(torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0)))
            put(S, "b_qkv", f32(torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0)))
            put(S, "w_ao", bf(a.output.dense.weight)); put(S, "b_ao", f32(a.output.dense.bias))
            put(S, "ln_a_g", f32(a.output.LayerNorm.weight)); put(S, "ln_a_b", f32(a.output.LayerNorm.bias))
            if L.has_cross_attention:
                c = L.crossattention
                put(S, "w_cq", bf(c.self.query.weight)); put(S, "b_cq", f32(c.self.query.bias))
                put(S, "w_co", bf(c.output.dense.weight)); put(S, "b_co", f32(c.output.dense.bias))
                put(S, "ln_c_g", f32(c.output.LayerNorm.weight)); put(S, "ln_c_b", f32(c.output.LayerNorm.bias))
                ckv_w += [c.self.key.weight, c.self.value.weight]
                ckv_b += [c.self.key.bias, c.self.value.bias]
            put(S, "w_fq1", bf(L.intermediate_query.dense.weight)); put(S, "b_fq1", f32(L.intermediate_query.dense.bias))
            put(S, "w_fq2", bf(L.output_query.dense.weight)); put(S, "b_fq2", f32(L.output_query.dense.bias))
            put(S, "ln_fq_g", f32(L.output_query.LayerNorm.weight)); put(S, "ln_fq_b", f32(L.output_query.LayerNorm.bias))
            put(S, "w_ft1", bf(L.intermediate.dense.weight)); put(S, "b_ft1", f32(L.intermediate.dense.bias))
            put(S, "w_ft2", bf(L.output.dense.weight)); put(S, "b_ft2", f32(L.output.dense.bias))
            put(S, "ln_ft_g", f32(L.output.LayerNorm.weight)); put(S, "ln_ft_b", f32(L.output.LayerNorm.bias))
        put(W, "w_ckv", bf(torch.cat(ckv_w, 0)))
        put(W, "b_ckv", f32(torch.cat(ckv_b, 0)))
        if proj is not None:
            pw, pb = _packed_proj(proj)
            put(W, "w_proj", pw)
            put(W, "b_proj", pb)
        self._pack = (ver, keep, W)
        self._pack_applied = set()   # handles that have seen this pack
        return W

    # ------------------------------------------------------------------------------------------------------- forward
    def prepare(self, input_ids=None, attention_mask=None, position_ids=None, query_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, llm_proj: nn.Linear = None, need_last_hidden: bool = True,
                skip_dead_text_ffn: bool = False, llm_scatter=None, llm_out=None, enc_ready=None):
        """``llm_scatter = (view [bs, F, Nq, D] into inputs_embeds)``: llm_proj writes each frame's tokens into that strided
        view (4-D TMA store, see ``mraudio_b200/prompt.py``) instead of a dense ``[rows*Nq, D]`` tensor.
        ``llm_out``: caller-owned bf16 ``[rows*Nq, D]`` buffer for the projected tokens (streaming callers that reuse their
        buffers: no allocation per call, see ``HostPipeline``).  ``enc_ready``: a recorded ``torch.cuda.Event`` after which
        ``encoder_hidden_states`` is valid (it is being copied in on another stream): only the kernels that read it wait.

        Validate the arguments of one ``Qformer.bert(...)`` call and stage everything the C-ABI needs (handle, packed
        weights, io struct, workspace, output tensors) without launching.  ``launch_prepared`` enqueues one or two such
        calls (two = both modalities in lockstep, grouped GEMM launches)."""
        if query_embeds is None or encoder_hidden_states is None:
            raise ValueError("BertModelB200 needs query_embeds and encoder_hidden_states (the Q-Former call of "
                             "models/xinstructblip.py:286-293)")
        if position_ids is not None:
            raise NotImplementedError("custom position_ids are not used by the reference path")
        enc = encoder_hidden_states
        if not enc.is_cuda:
            raise _lib.MraError("BertModelB200.forward needs CUDA tensors: mraudio_b200 has no CPU fallback")
        cfg = self.config
        dev = enc.device
        rows, Nk, Wd = enc.shape
        if Wd != cfg.encoder_width:
            raise ValueError(f"encoder_hidden_states width {Wd} != config.encoder_width {cfg.encoder_width}")
        Nq = query_embeds.shape[1]
        if Nq != cfg.query_length:
            raise ValueError(f"query_embeds has {Nq} tokens, config.query_length is {cfg.query_length}")
        if enc_ready is not None and (enc.dtype != torch.bfloat16 or not enc.is_contiguous()):
            torch.cuda.current_stream(dev).wait_event(enc_ready)    # the conversion below reads it: wait here instead
            enc_ready = None
        enc_b = enc.to(torch.bfloat16).contiguous()
        q_rows = query_embeds.shape[0]
        # ``query_tokens.expand(bs,-1,-1).repeat(F,1,1)`` (:229,289) materialises identical rows: detect the broadcast
        if q_rows != 1 and query_embeds.stride(0) == 0:
            query_embeds, q_rows = query_embeds[:1], 1
        qe = query_embeds.detach().float().contiguous()
        if q_rows not in (1, rows):
            raise ValueError(f"query_embeds batch {q_rows} does not match encoder rows {rows}")
        T = 0
        ids = tmask = emask = None
        if input_ids is not None:
            if input_ids.shape[0] != rows:
                raise ValueError(f"input_ids batch {input_ids.shape[0]} does not match encoder rows {rows}")
            T = input_ids.shape[1]
            ids = input_ids.to(device=dev, dtype=torch.int32).contiguous()
        # masks go to the device as they are (no host-side inspection: that would synchronise the stream)
        if attention_mask is not None:
            if attention_mask.shape != (rows, Nq + T):
                raise ValueError(f"attention_mask shape {tuple(attention_mask.shape)} != {(rows, Nq + T)}")
            tmask = attention_mask.to(device=dev, dtype=torch.int32).contiguous()
        if encoder_attention_mask is not None:
            if encoder_attention_mask.shape != (rows, Nk):
                raise ValueError(f"encoder_attention_mask shape {tuple(encoder_attention_mask.shape)} != {(rows, Nk)}")
            emask = encoder_attention_mask.to(device=dev, dtype=torch.int32).contiguous()

        llm_dim = llm_proj.weight.shape[0] if llm_proj is not None else 0
        h = self._handle(llm_dim)
        W = self._packed(llm_proj)
        if llm_dim not in self._pack_applied:
            check(lib.mra_qformer_set_weights(h, C.byref(W)))
            self._pack_applied.add(llm_dim)

        flags = _lib.FWD_SKIP_DEAD_TEXT_FFN if skip_dead_text_ffn else 0
        need = lib.mra_qformer_workspace_bytes(h, rows, T, Nk, flags)
        if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
            self._workspace = torch.empty(need, device=dev, dtype=torch.uint8)
        last_hidden = torch.empty(rows, Nq + T, cfg.hidden_size, device=dev, dtype=torch.float32) if need_last_hidden else None
        scatter = {}
        if llm_scatter is not None:
            v = llm_scatter
            if llm_proj is None or v.dtype != torch.bfloat16 or v.dim() != 4 or v.shape[2:] != (Nq, llm_dim) or \
                    v.shape[0] * v.shape[1] != rows or v.stride(3) != 1:
                raise ValueError(f"llm_scatter must be a bf16 [bs, F, {Nq}, {llm_dim}] view with bs*F == {rows}")
            llm_out, llm_view = v, v
            scatter = dict(llm_frames=v.shape[1], llm_ld=v.stride(2), llm_frame_stride=v.stride(1), llm_video_stride=v.stride(0))
        else:
            if llm_out is not None:
                if llm_proj is None or llm_out.dtype != torch.bfloat16 or llm_out.device != dev or not llm_out.is_contiguous() or \
                        llm_out.numel() != rows * Nq * llm_dim:
                    raise ValueError(f"llm_out must be a contiguous bf16 buffer of {rows * Nq} x {llm_dim} elements on {dev}")
                llm_out = llm_out.view(rows * Nq, llm_dim)
            else:
                llm_out = torch.empty(rows * Nq, llm_dim, device=dev, dtype=torch.bfloat16) if llm_proj is not None else None
            llm_view = llm_out.view(rows, Nq, llm_dim) if llm_out is not None else None
        io = _lib.QFormerIO(enc=enc_b.data_ptr(), input_ids=_lib.ptr(ids), attn_mask=_lib.ptr(tmask), enc_mask=_lib.ptr(emask),
                            query_embeds=qe.data_ptr(), q_rows=q_rows, rows=rows, T=T, Nk=Nk, flags=flags,
                            last_hidden=_lib.ptr(last_hidden), llm_out=_lib.ptr(llm_out),
                            enc_ready=enc_ready.cuda_event if enc_ready is not None else None, **scatter)
        out = QFormerOutput(last_hidden_state=last_hidden, llm_inputs=llm_view)
        return _Prepared(self, h, io, self._workspace, out, (enc_b, qe, ids, tmask, emask))   # (enc_ready: owned by the caller)

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, query_embeds=None,
                encoder_hidden_states=None, encoder_attention_mask=None, return_dict=True, llm_proj: nn.Linear = None,
                need_last_hidden: bool = True, skip_dead_text_ffn: bool = False, **unused):
        prep = self.prepare(input_ids, attention_mask, position_ids, query_embeds, encoder_hidden_states,
                            encoder_attention_mask, llm_proj, need_last_hidden, skip_dead_text_ffn)
        launch_prepared([prep])
        return prep.out if return_dict else (prep.out.last_hidden_state,)


class _Prepared:
    __slots__ = ("bert", "handle", "io", "workspace", "out", "keep")

    def __init__(self, bert, handle, io, workspace, out, keep):
        self.bert, self.handle, self.io, self.workspace, self.out, self.keep = bert, handle, io, workspace, out, keep


def launch_prepared(preps) -> int:
    """Enqueue 1 or 2 prepared Q-Former calls on the current stream (2 = lockstep with grouped GEMM launches, see
    ``mra_qformer_forward_multi``).  Returns the number of kernels launched."""
    n = len(preps)
    if n == 1:
        p = preps[0]
        check(lib.mra_qformer_forward(p.handle, C.byref(p.io), p.workspace.data_ptr(), p.workspace.numel(), current_stream()))
    else:
        hs = (C.c_void_p * n)(*[p.handle for p in preps])
        ios = (C.POINTER(_lib.QFormerIO) * n)(*[C.pointer(p.io) for p in preps])
        wss = (C.c_void_p * n)(*[p.workspace.data_ptr() for p in preps])
        wsb = (C.c_size_t * n)(*[p.workspace.numel() for p in preps])
        check(lib.mra_qformer_forward_multi(n, hs, ios, wss, wsb, current_stream()))
    launches = 0
    stream = torch.cuda.current_stream()
    for p in preps:
        p.bert.last_launches = lib.mra_qformer_last_launch_count(p.handle)
        launches += p.bert.last_launches
        for t in p.keep:   # the tensors handed to the asynchronous launch must outlive it on this stream
            if t is not None:
                t.record_stream(stream)
    return launches


class BertLMHeadModel(nn.Module):
    """``lavis.models.blip2_models.Qformer.BertLMHeadModel`` stand-in (models/xinstructblip.py:15,623).  Only ``.bert`` is
    on the hot path; ``.cls`` exists so that ``Qformer.cls = None`` (:135) works and LM-head keys are reported as
    unexpected by ``load_state_dict(strict=False)`` just as they are dropped upstream."""

    def __init__(self, config: BertConfig):
        super().__init__()
        self.config = config
        self.bert = BertModelB200(config)
        self.cls = None

    def resize_token_embeddings(self, new_num_tokens: int):
        # models/xinstructblip.py:134 -- grow / shrink the word-embedding table, keeping the common rows
        emb = self.bert.embeddings.word_embeddings
        old = emb.weight.data
        if new_num_tokens == old.shape[0]:
            return emb
        new = nn.Embedding(new_num_tokens, old.shape[1], padding_idx=emb.padding_idx).to(old.device, old.dtype)
        new.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        n = min(new_num_tokens, old.shape[0])
        new.weight.data[:n] = old[:n]
        self.bert.embeddings.word_embeddings = new
        self.config.vocab_size = new_num_tokens
        for hnd in self.bert._handles.values():
            lib.mra_qformer_destroy(hnd)
        self.bert._handles = {}
        self.bert._pack = None
        self.bert._pack_applied = set()
        return new

    def forward(self, *a, **k):
        raise NotImplementedError("only Qformer.bert(...) is on the mrAudio hot path (models/xinstructblip.py:286)")
