"""Host-side mirror of the Q-Former section of ``models/xinstructblip.py`` (the part of ``XInstructBLIP`` between the
frozen encoders and the LLM prompt assembly), backed by libmraudio_b200.

Reference surface kept (file:line under /root/reference/models/xinstructblip.py):

* ``init_Qformer(num_query_token, modality_width, cross_attention_freq=2, pretrained_qformer=None, load_attention=False,
  load_qformer_type="")`` -> ``(Qformer, query_tokens)``                                        (:615-655)
* ``init_ln(num_features, load_ln_path, load_ln_type)`` -> ``LayerNorm`` (fp32-upcast)            (:679-704, :822-828)
* ``init_vicuna_projection(input_size, output_size, load_projection_path, load_projection_type)`` (:707-735)
* module attributes ``{modality}_Qformer``, ``{modality}_query_tokens``, ``{modality}_ln``, ``{modality}_llm_proj``,
  ``num_query_token``, ``modalities`` (:119-189) and the freezing loop (:196-204)
* ``encode_modalities`` = the arithmetic of ``generate`` :228-305 / ``forward`` :406-477 for given encoder outputs:
  query-token expand, mask build, frame fold + batch-major reorder, ``Qformer.bert``, slice, ``llm_proj``, reshape to
  ``[bs, F*32, D]``, ``atts_llm`` -- including the reference's frame-major text tiling for bs > 1 (:287-289).

* ``encode_modalities(..., prompt=PromptPieces)`` continues through the LLM-prompt assembly of :342-386 / :544-594
  (``mraudio_b200/prompt.py``) and returns ``(inputs_embeds, attention_mask)``.

Everything out of scope (ViT / BEATs encoders, tokenizers, the LLM and its embedding table) is taken as input: the
encoders' outputs ("cached features"), BERT ``input_ids`` / ``attention_mask`` of the prompt, embeddings of LLM tokens.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import nn

from . import _lib, ops
from . import prompt as prompt_mod
from .qformer import BertConfig, BertLMHeadModel, LLMProjB200, launch_prepared

Features = Union[torch.Tensor, Sequence[torch.Tensor]]


class LayerNorm(nn.LayerNorm):
    """``{modality}_ln`` (models/xinstructblip.py:822-828): LayerNorm computed in fp32 and cast back.  The CUDA kernel
    fuses it with the frame fold; ``forward`` on a plain tensor keeps the reference semantics (output dtype == input)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.MraError("LayerNorm needs CUDA tensors (no CPU fallback)")
        shp = x.shape
        x4 = x.reshape(1, 1, -1, shp[-1]).contiguous()
        if x4.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            x4 = x4.float()
        y = ops.modality_layernorm(x4, self.weight.detach().float().contiguous(), self.bias.detach().float().contiguous(),
                                   self.eps)
        return y.reshape(shp).to(x.dtype)


def _load_ckpt(url_or_filename: str):
    if not os.path.isfile(url_or_filename):
        raise RuntimeError("checkpoint url or path is invalid")  # same error as the reference (:645, :697, :725)
    ckpt = torch.load(url_or_filename, map_location="cpu")
    return ckpt["model"] if "model" in ckpt else ckpt


class XInstructBLIPQFormers(nn.Module):
    """The Q-Former / projection sub-graph of ``XInstructBLIP`` with the reference's attribute names."""

    def __init__(self, modalities: Sequence[str] = ("audio", "video"), encoder_num_features: Optional[Dict[str, int]] = None,
                 llm_hidden_size: int = 4096, num_query_token: int = 32, tokenizer_len: int = 30523,
                 pretrained_qformers: Optional[Dict[str, str]] = None, num_hidden_layers: int = 12,
                 freeze: bool = True):
        super().__init__()
        self.modalities = list(modalities)                                  # :71
        self.num_query_token = num_query_token                              # :120
        self.llm_hidden_size = llm_hidden_size                              # :161
        widths = dict(encoder_num_features or {"video": 1408, "audio": 768})
        pretrained_qformers = pretrained_qformers or {}
        for modality in self.modalities:
            pre = pretrained_qformers.get(modality)
            setattr(self, f"{modality}_ln", self.init_ln(widths[modality], load_ln_path=pre, load_ln_type=modality))
            qformer, query_tokens = self.init_Qformer(num_query_token, widths[modality], pretrained_qformer=pre,
                                                      load_attention=True, load_qformer_type=modality,
                                                      num_hidden_layers=num_hidden_layers)
            qformer.resize_token_embeddings(tokenizer_len)                  # :134
            qformer.cls = None                                              # :135
            setattr(self, f"{modality}_Qformer", qformer)
            setattr(self, f"{modality}_query_tokens", query_tokens)
            proj = self.init_vicuna_projection(qformer.config.hidden_size, llm_hidden_size, load_projection_path=pre,
                                               load_projection_type=modality)
            setattr(self, f"{modality}_llm_proj", proj)
        if freeze:
            self.freeze_qformers()
        self.lockstep_modalities = True     # both Q-Formers in one lockstep call with grouped GEMM launches
        self.last_hidden_states = {}
        self.last_launches = 0

    def freeze_qformers(self, frozen: bool = True):
        """:196-204 -- as shipped the Q-Formers, query tokens, LNs and projections are frozen; ``frozen=False`` is the
        north-star finetune configuration (Q-Former/projection trainable, encoders and LLM frozen)."""
        for modality in self.modalities:
            for _, p in getattr(self, f"{modality}_ln").named_parameters():
                p.requires_grad = False
            getattr(self, f"{modality}_query_tokens").requires_grad = not frozen
            for _, p in getattr(self, f"{modality}_Qformer").named_parameters():
                p.requires_grad = not frozen
            for _, p in getattr(self, f"{modality}_llm_proj").named_parameters():
                p.requires_grad = not frozen

    # ------------------------------------------------------------------------------------------------ constructors
    @classmethod
    def init_Qformer(cls, num_query_token, modality_width, cross_attention_freq=2, pretrained_qformer=None,
                     load_attention=False, load_qformer_type="", num_hidden_layers=12):
        encoder_config = BertConfig.from_pretrained("bert-base-uncased")
        encoder_config.encoder_width = modality_width
        encoder_config.add_cross_attention = True
        encoder_config.cross_attention_freq = cross_attention_freq
        encoder_config.query_length = num_query_token
        encoder_config.num_hidden_layers = num_hidden_layers
        encoder_config.vocab_size += 1  # [DEC]
        Qformer = BertLMHeadModel(config=encoder_config)
        query_tokens = nn.Parameter(torch.zeros(1, num_query_token, encoder_config.hidden_size))
        query_tokens.data.normal_(mean=0.0, std=encoder_config.initializer_range)
        if pretrained_qformer:
            checkpoint = _load_ckpt(pretrained_qformer)
            if load_qformer_type:
                load_qformer_type = f"{load_qformer_type}_"
            loaded = {}
            for k in checkpoint.keys():
                if load_qformer_type + "Qformer." in k:
                    if not load_attention and "attention" in k:
                        continue
                    loaded[".".join(k.split(".")[1:])] = checkpoint[k]
            Qformer.load_state_dict(loaded, strict=False)
            query_tokens.data = checkpoint[load_qformer_type + "query_tokens"]
        return Qformer, query_tokens

    @classmethod
    def init_ln(cls, num_features, load_ln_path=False, load_ln_type=""):
        ln = LayerNorm(num_features)
        if load_ln_path and load_ln_type:
            checkpoint = _load_ckpt(load_ln_path)
            key = f"{load_ln_type}_ln" if "vision" not in load_ln_type else "ln_vision"
            loaded = {".".join(k.split(".")[1:]): v for k, v in checkpoint.items() if key in k}
            ln.load_state_dict(loaded, strict=False)
        return ln

    @classmethod
    def init_vicuna_projection(cls, input_size, output_size, load_projection_path=False, load_projection_type="",
                               projection_key=None):
        proj = LLMProjB200(input_size, output_size)
        if load_projection_path:
            checkpoint = _load_ckpt(load_projection_path)
            if load_projection_type:
                load_projection_type = f"{load_projection_type}_"
            loaded = {}
            for k in checkpoint.keys():
                if projection_key:
                    if projection_key in k:
                        loaded[".".join(k.split(".")[1:])] = checkpoint[k]
                elif load_projection_type + "llm_proj." in k:
                    loaded[".".join(k.split(".")[1:])] = checkpoint[k]
            proj.load_state_dict(loaded, strict=False)
        return proj

    # ------------------------------------------------------------------------------------------------------ hot path
    def fold_frames(self, modality: str, feats: Features, apply_ln: bool) -> torch.Tensor:
        """:262-285.  ``feats`` is either the reference's per-frame list (F tensors ``[bs, Nk, W]``, i.e. ``embeds[m]``
        before ``torch.cat``) or a stacked ``[bs, F, Nk, W]`` tensor (cached-feature layout).  Returns batch-major rows
        ``[bs*F, Nk, W]`` in bf16.  With ``apply_ln`` the inputs are raw encoder outputs and ``{modality}_ln`` is fused
        into the same pass (one read, one bf16 write); without it they are already LayerNorm'ed."""
        ln = getattr(self, f"{modality}_ln")
        if isinstance(feats, (list, tuple)):
            x = torch.stack(list(feats), 0)       # [F, bs, Nk, W] frame-major, what torch.cat(embeds) holds
            frame_major = True
            Fr, bs = x.shape[:2]
        else:
            x, frame_major = feats, False
            bs, Fr = x.shape[:2]
        if not x.is_cuda:
            raise _lib.MraError("encoder features must be CUDA tensors (no CPU fallback)")
        if apply_ln:
            return ops.modality_layernorm(x.contiguous(), ln.weight.detach().float().contiguous(),
                                          ln.bias.detach().float().contiguous(), ln.eps, frame_major=frame_major)
        if frame_major:
            x = x.transpose(0, 1)
        return x.reshape(bs * Fr, *x.shape[2:]).to(torch.bfloat16).contiguous()

    def encode_modalities(self, feats: Dict[str, Features], input_ids: torch.Tensor, attention_mask: torch.Tensor,
                          apply_ln: bool = False, match_reference_text_tiling: bool = True,
                          need_last_hidden: bool = False, prompt: Optional["prompt_mod.PromptPieces"] = None,
                          scatter_epilogue: bool = True, out: Optional[Dict[str, torch.Tensor]] = None,
                          ready: Optional[Dict[str, torch.cuda.Event]] = None):
        """Returns ``(inputs_llm, atts_llm)`` dicts exactly as :296-306 builds them.

        input_ids / attention_mask: ``text_Qformer.input_ids`` / ``.attention_mask`` ``[bs, T]`` (:233-239).

        With ``prompt`` (the LLM-token pieces of :320-386) the call continues through the prompt assembly and returns
        ``(inputs_embeds [bs, L, D], attention_mask [bs, L])`` -- what :388-392 hands to ``llm_model.generate``: the
        llm_proj epilogue writes the query tokens straight into their slots (``scatter_epilogue``; off = dense outputs +
        copy, kept for the parity test) and one more launch copies the text pieces.

        ``out``: per-modality caller-owned bf16 buffers ``[bs, F*32, D]`` the projected tokens are written into (streaming
        callers; the returned ``inputs_llm`` are views of them).  ``ready``: per-modality recorded events after which that
        modality's features are valid (copied in on another stream): only the kernels that read them wait.
        """
        inputs_llm, atts_llm = {}, {}
        todo = [m for m in self.modalities if m in feats]
        lay = embeds = None
        if prompt is not None:
            bs0 = input_ids.shape[0]
            f0 = feats[todo[0]]
            frames = len(f0) if isinstance(f0, (list, tuple)) else f0.shape[1]
            prompt = prompt_mod.normalise_pieces(prompt, input_ids.device)
            lay = prompt_mod.PromptLayout.build(prompt, bs0, frames, self.num_query_token, todo)
            embeds = torch.empty(bs0, lay.L, lay.D, device=input_ids.device, dtype=torch.bfloat16)
            scatter_epilogue = scatter_epilogue and lay.uniform and self.num_query_token == 32
        preps, shapes = [], []
        extra = 0
        for modality in todo:
            ev = ready.get(modality) if ready is not None else None
            if ev is not None and (apply_ln or isinstance(feats[modality], (list, tuple))):
                torch.cuda.current_stream(input_ids.device).wait_event(ev)    # the fold below launches kernels that read them
                ev = None
            enc = self.fold_frames(modality, feats[modality], apply_ln)
            extra += 1 if apply_ln else 0
            bs = input_ids.shape[0]
            num = enc.shape[0] // bs
            if match_reference_text_tiling:
                ids = input_ids.repeat(num, 1)                                    # :287 (frame-major tiling)
                tmask = attention_mask.repeat(num, 1)
            else:
                ids = input_ids.repeat_interleave(num, 0)
                tmask = attention_mask.repeat_interleave(num, 0)
            query_tokens = getattr(self, f"{modality}_query_tokens")
            q_atts = torch.ones(enc.shape[0], self.num_query_token, dtype=tmask.dtype, device=tmask.device)   # :246
            qformer = getattr(self, f"{modality}_Qformer")
            proj = getattr(self, f"{modality}_llm_proj")
            preps.append(qformer.bert.prepare(ids, attention_mask=torch.cat([q_atts, tmask], 1), query_embeds=query_tokens,
                                              encoder_hidden_states=enc, encoder_attention_mask=None, llm_proj=proj,
                                              need_last_hidden=need_last_hidden, skip_dead_text_ffn=not need_last_hidden,
                                              llm_scatter=prompt_mod.query_slot_view(embeds, lay, modality)
                                              if lay is not None and scatter_epilogue else None,
                                              llm_out=out[modality] if out is not None and lay is None else None,
                                              enc_ready=ev))
            shapes.append((bs, num))
        # Both Q-Formers share the layer geometry: run them in lockstep, every Linear as ONE grouped GEMM launch over
        # (video queries, video text, audio queries, audio text) -- see mra_qformer_forward_multi.
        launches = 0
        if self.lockstep_modalities and len(preps) == 2:
            launches = launch_prepared(preps)
        else:
            for p in preps:
                launches += launch_prepared([p])
        self.last_launches = launches + extra
        if lay is not None:
            dense = None if scatter_epilogue else {m: p.out.llm_inputs.view(lay.bs, lay.frames, self.num_query_token, lay.D)
                                                   for m, p in zip(todo, preps)}
            self.last_launches += prompt_mod.copy_pieces(embeds, prompt, lay, dense)
            self.last_prompt_layout = lay
            return embeds, prompt_mod.attention_mask(prompt, lay, embeds.device)
        for modality, p, (bs, num) in zip(todo, preps, shapes):
            y = p.out.llm_inputs                                                   # [bs*num, 32, D]
            inputs_llm[modality] = y.reshape(bs, num, self.num_query_token, -1).view(bs, num * self.num_query_token, -1)
            atts_llm[modality] = torch.ones(inputs_llm[modality].size()[:-1], dtype=torch.long, device=y.device)  # :306
            if need_last_hidden:
                self.last_hidden_states[modality] = p.out.last_hidden_state
        return inputs_llm, atts_llm

    def host_pipeline(self, bs: int, frames: int, tokens: Dict[str, int], text_len: int, slots: int = 2) -> "HostPipeline":
        """Streaming host entry: pinned host features in, pinned host ``inputs_llm`` out (see ``HostPipeline``)."""
        return HostPipeline(self, bs, frames, tokens, text_len, slots)


class HostPipeline:
    """Double-buffered host <-> device pipeline around ``encode_modalities`` for callers whose encoder features live in
    host memory (cached features on disk / produced by another process).  ``submit`` enqueues, on three streams,
    H2D copy of the batch -> both Q-Formers + projections -> D2H copy of ``inputs_llm`` into this slot's pinned output
    buffers, and returns the slot; nothing blocks the host until ``drain`` / ``slot.out_done.synchronize()``."""

    class _Slot:
        pass

    def __init__(self, model: XInstructBLIPQFormers, bs: int, frames: int, tokens: Dict[str, int], text_len: int, slots: int):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.MraError("HostPipeline needs the model on a CUDA device (no CPU fallback)")
        self.model, self.dev = model, dev
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.slots: List[HostPipeline._Slot] = []
        self.i = 0
        nq, D = model.num_query_token, model.llm_hidden_size
        for _ in range(slots):
            sl = HostPipeline._Slot()
            sl.feats, sl.out_dev, sl.out_host = {}, {}, {}
            for m in model.modalities:
                W = getattr(model, f"{m}_Qformer").config.encoder_width
                sl.feats[m] = torch.empty(bs, frames, tokens[m], W, device=dev, dtype=torch.bfloat16)
                sl.out_dev[m] = torch.empty(bs, frames * nq, D, device=dev, dtype=torch.bfloat16)
                sl.out_host[m] = torch.empty(bs, frames * nq, D, dtype=torch.bfloat16).pin_memory()
            sl.ids = torch.empty(bs, text_len, device=dev, dtype=torch.long)
            sl.mask = torch.empty(bs, text_len, device=dev, dtype=torch.long)
            sl.in_done, sl.compute_done, sl.out_done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            sl.text_done = torch.cuda.Event()
            sl.feat_done = {m: torch.cuda.Event() for m in model.modalities}
            sl.used = False
            self.slots.append(sl)
        s0 = self.slots[0]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in s0.feats.values()) + \
            s0.ids.numel() * 8 + s0.mask.numel() * 8
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in s0.out_host.values())

    def submit(self, host_feats: Dict[str, torch.Tensor], input_ids: torch.Tensor, attention_mask: torch.Tensor):
        sl = self.slots[self.i % len(self.slots)]
        self.i += 1
        main = torch.cuda.current_stream(self.dev)
        if sl.used:
            self.copy_in.wait_event(sl.compute_done)   # the previous batch in this slot has been consumed
            main.wait_event(sl.out_done)               # ... and its projected tokens have left the slot's device buffers
        with torch.cuda.stream(self.copy_in):
            # prompt tokens first (tiny), then the modalities in the order the forward reads them; one event each, so that the
            # first Q-Former's cross-K/V projection starts while the second one's features are still on the bus
            sl.ids.copy_(input_ids, non_blocking=True)
            sl.mask.copy_(attention_mask, non_blocking=True)
            sl.text_done.record(self.copy_in)
            for m in self.model.modalities:
                if m in host_feats:
                    sl.feats[m].copy_(host_feats[m], non_blocking=True)
                    sl.feat_done[m].record(self.copy_in)
            sl.in_done.record(self.copy_in)
        main.wait_event(sl.text_done)
        ready = {m: sl.feat_done[m] for m in self.model.modalities if m in host_feats}
        with torch.no_grad():
            # the projections write into this slot's own device buffers: no output allocation per batch (fresh tensors handed
            # to another stream keep their blocks reserved until that stream catches up, and a host that runs ahead of the
            # device then sends the caching allocator to cudaMalloc in the middle of the stream: stalls of 10-80 ms per call)
            inputs_llm, _ = self.model.encode_modalities({m: sl.feats[m] for m in ready}, sl.ids, sl.mask, out=sl.out_dev, ready=ready)
        sl.compute_done.record(main)
        self.copy_out.wait_event(sl.compute_done)
        with torch.cuda.stream(self.copy_out):
            for m, y in inputs_llm.items():
                sl.out_host[m].copy_(y, non_blocking=True)
            sl.out_done.record(self.copy_out)
        sl.used = True
        return sl

    def drain(self):
        """Make the current stream wait for every outstanding D2H copy (so an event recorded next covers them)."""
        main = torch.cuda.current_stream(self.dev)
        for sl in self.slots:
            if sl.used:
                main.wait_event(sl.out_done)
