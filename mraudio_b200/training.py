"""Fine-tuning of the Q-Former / query tokens / llm_proj on cached encoder features (LLM and encoders frozen).

Mirrors the slice of ``utils/trainer.py`` that touches the hot path (file:line under /root/reference):

* ``Trainer.train_epoch`` inner loop ``:124-140`` -- forward, ``loss / accum_grad_iters``, backward, optimizer step every
  ``accum_grad_iters`` iterations, ``zero_grad``;
* ``optim.Adam(model.parameters(), lr=3e-4)`` ``:65`` and ``LinearWarmupCosineLRScheduler(max_epoch, min_lr=0,
  init_lr=3e-4, warmup_steps=1000, warmup_start_lr=1e-8)`` ``:66`` stepped per iteration ``:127``;
* ``DistributedDataParallel`` ``:69``: gradients are averaged over ranks -- here by ONE NCCL all-reduce per optimizer
  step over the flat gradient buffer of each modality (the reference all-reduces on every backward, also on
  non-stepping accumulation iterations);
* ``_save_checkpoint`` ``:184-210``: only ``requires_grad`` parameters + optimizer state + epoch.

As shipped the reference freezes these parameters (``models/xinstructblip.py:196-204``) and trains LoRA adapters of the
LLM; un-freezing them is the training configuration BASELINE.json's north_star names.  bf16 tensor-core math with fp32
master weights replaces the reference's fp16 autocast + GradScaler (no loss scaling is needed with bf16 exponents).

Memory layout: all trainable tensors of one modality live in ONE flat fp32 buffer in the packed order the C-ABI wants
(q,k,v stacked; cross k,v of all layers stacked); the ``nn.Parameter``s of the module become views into it, ``.grad``s
are views into a second flat buffer, Adam state two more.  One cast kernel refreshes the bf16 operand copy.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import check, current_stream, lib
from .qformer import BertLMHeadModel

_ALIGN = 64  # elements


def _layer_segments(H, I, W, cross):
    seg = [("w_qkv", (3 * H, H)), ("b_qkv", (3 * H,)), ("w_ao", (H, H)), ("b_ao", (H,)), ("ln_a_g", (H,)), ("ln_a_b", (H,))]
    if cross:
        seg += [("w_cq", (H, H)), ("b_cq", (H,)), ("w_co", (H, H)), ("b_co", (H,)), ("ln_c_g", (H,)), ("ln_c_b", (H,))]
    seg += [("w_fq1", (I, H)), ("b_fq1", (I,)), ("w_fq2", (H, I)), ("b_fq2", (H,)), ("ln_fq_g", (H,)), ("ln_fq_b", (H,)),
            ("w_ft1", (I, H)), ("b_ft1", (I,)), ("w_ft2", (H, I)), ("b_ft2", (H,)), ("ln_ft_g", (H,)), ("ln_ft_b", (H,))]
    return seg


class TrainableQFormer:
    """Flat-buffer training state of one modality: ``{modality}_Qformer`` + ``{modality}_query_tokens`` +
    ``{modality}_llm_proj``.  After construction the module's parameters alias ``self.flat`` (fp32 master weights)."""

    def __init__(self, qformer: BertLMHeadModel, query_tokens: nn.Parameter, llm_proj: nn.Linear):
        bert = qformer.bert
        dev = query_tokens.device
        if dev.type != "cuda":
            raise _lib.MraError("TrainableQFormer needs the module on a CUDA device (no CPU fallback)")
        self.qformer, self.bert, self.query_tokens, self.llm_proj = qformer, bert, query_tokens, llm_proj
        cfg = bert.config
        self.cfg = cfg
        H, I, W, D = cfg.hidden_size, cfg.intermediate_size, cfg.encoder_width, llm_proj.weight.shape[0]
        self.H, self.I, self.W, self.D = H, I, W, D
        layers = bert.encoder.layer
        self.cross = [bool(L.has_cross_attention) for L in layers]
        nc = sum(self.cross)
        # ---- segment table (name -> offset, shape) in the packed order
        segs: List[Tuple[str, tuple]] = [("word_emb", tuple(bert.embeddings.word_embeddings.weight.shape)),
                                         ("pos_emb", tuple(bert.embeddings.position_embeddings.weight.shape)),
                                         ("ln_e_g", (H,)), ("ln_e_b", (H,)), ("w_ckv", (nc * 2 * H, W)), ("b_ckv", (nc * 2 * H,)),
                                         ("w_proj", (D, H)), ("b_proj", (D,)), ("query_tokens", tuple(query_tokens.shape))]
        for l, c in enumerate(self.cross):
            segs += [(f"L{l}.{n}", shp) for n, shp in _layer_segments(H, I, W, c)]
        self.seg: Dict[str, Tuple[int, tuple]] = {}
        off = 0
        for name, shp in segs:
            self.seg[name] = (off, shp)
            off += (math.prod(shp) + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat16 = torch.zeros(off, device=dev, dtype=torch.bfloat16)
        # ---- re-home every parameter into the flat buffer (and its .grad into the flat gradient buffer)
        e = bert.embeddings
        self._bind(e.word_embeddings.weight, "word_emb")
        self._bind(e.position_embeddings.weight, "pos_emb")
        self._bind(e.LayerNorm.weight, "ln_e_g")
        self._bind(e.LayerNorm.bias, "ln_e_b")
        self._bind(llm_proj.weight, "w_proj")
        self._bind(llm_proj.bias, "b_proj")
        self._bind(query_tokens, "query_tokens")
        slot = 0
        for l, L in enumerate(layers):
            p = f"L{l}."
            a = L.attention
            for i, n in enumerate(("query", "key", "value")):
                self._bind(getattr(a.self, n).weight, p + "w_qkv", i * H, H)
                self._bind(getattr(a.self, n).bias, p + "b_qkv", i * H, H)
            self._bind(a.output.dense.weight, p + "w_ao"); self._bind(a.output.dense.bias, p + "b_ao")
            self._bind(a.output.LayerNorm.weight, p + "ln_a_g"); self._bind(a.output.LayerNorm.bias, p + "ln_a_b")
            if self.cross[l]:
                c = L.crossattention
                self._bind(c.self.query.weight, p + "w_cq"); self._bind(c.self.query.bias, p + "b_cq")
                self._bind(c.output.dense.weight, p + "w_co"); self._bind(c.output.dense.bias, p + "b_co")
                self._bind(c.output.LayerNorm.weight, p + "ln_c_g"); self._bind(c.output.LayerNorm.bias, p + "ln_c_b")
                self._bind(c.self.key.weight, "w_ckv", slot * 2 * H, H); self._bind(c.self.key.bias, "b_ckv", slot * 2 * H, H)
                self._bind(c.self.value.weight, "w_ckv", slot * 2 * H + H, H); self._bind(c.self.value.bias, "b_ckv", slot * 2 * H + H, H)
                slot += 1
            self._bind(L.intermediate_query.dense.weight, p + "w_fq1"); self._bind(L.intermediate_query.dense.bias, p + "b_fq1")
            self._bind(L.output_query.dense.weight, p + "w_fq2"); self._bind(L.output_query.dense.bias, p + "b_fq2")
            self._bind(L.output_query.LayerNorm.weight, p + "ln_fq_g"); self._bind(L.output_query.LayerNorm.bias, p + "ln_fq_b")
            self._bind(L.intermediate.dense.weight, p + "w_ft1"); self._bind(L.intermediate.dense.bias, p + "b_ft1")
            self._bind(L.output.dense.weight, p + "w_ft2"); self._bind(L.output.dense.bias, p + "b_ft2")
            self._bind(L.output.LayerNorm.weight, p + "ln_ft_g"); self._bind(L.output.LayerNorm.bias, p + "ln_ft_b")
        # (the dgrad GEMMs read the bf16 weights in place through MN-major descriptors: no transposed copies are kept)
        # The trainer owns its handle: sharing the inference handle of bert._handles would let a validation pass between
        # epochs (set_weights with an inference pack) re-point every later training step at a frozen snapshot.
        self._handle = bert._new_handle(D)
        self.version = 0                       # bumped whenever the master weights change behind autograd's back
        # ---- gradient buckets for the overlapped all-reduce, in the order in which the backward finalises them:
        #      the projection (first launches of the backward), then pairs of layers from the top, then what is final only at
        #      the end (embeddings, stacked cross K/V, query tokens).  Each entry: (event index | None = end of backward,
        #      [(lo, hi) ranges of the flat buffer]).  12 layers -> 8 buckets; the tail holds ~20 % of the elements (the word
        #      embedding table alone is 12.6 %) and is the only part whose exchange cannot overlap the backward.
        nl = len(layers)
        self.buckets: List[Tuple[Optional[int], List[Tuple[int, int]]]] = []
        proj_lo, proj_hi = self.seg["w_proj"][0], self.seg["query_tokens"][0]
        self.buckets.append((nl, [(proj_lo, proj_hi)]))
        hi = self.numel
        step = 2 if nl >= 4 else 1
        for b in range(nl - step, -1, -step):
            lo = self.seg[f"L{b}.w_qkv"][0]
            self.buckets.append((b, [(lo, hi)]))
            hi = lo
        if hi > self.seg["L0.w_qkv"][0]:                                    # odd layer count: the remaining bottom layer(s)
            self.buckets.append((0, [(self.seg["L0.w_qkv"][0], hi)]))
            hi = self.seg["L0.w_qkv"][0]
        self.buckets.append((None, [(0, proj_lo), (proj_hi, hi)]))
        self.layer_events = [torch.cuda.Event() for _ in range(nl + 1)]    # [nl] = projection gradients final
        for e in self.layer_events:
            e.record()                                                     # creates the underlying cudaEvent_t
        ev = (C.c_void_p * (nl + 1))(*[e.cuda_event for e in self.layer_events])
        check(lib.mra_qformer_backward_layer_events(self._handle, ev, nl + 1))
        # gradient exchange dtype: bf16 halves the bytes on NVLink (fp32 accumulation on every rank, bf16 on the wire --
        # the usual DDP compression hook); fp32 keeps the exchange exact (used by the equivalence checks)
        self.grad_comm_dtype = torch.bfloat16
        self.grad16 = None                      # bf16 staging / result of the exchange (allocated on first use)
        self._reduced_bf16 = False              # the last exchange left its result in grad16
        self.reduce_group = None
        self.backward_done = torch.cuda.Event()       # recorded on the backward's stream when mra_qformer_backward is enqueued
        self.backward_ran = False
        self._structs = None
        self._ws = self._bws = None
        self._saved = None
        self.step_count = 0
        self.refresh_operands()
        # inference through bert / llm_proj now reads these live buffers (see BertModelB200._packed / _packed_proj)
        bert._trainable = self
        bert._pack = None
        bert._pack_applied = set()
        llm_proj._mra_trainable = self
        llm_proj._mra_pack = None

    def __del__(self):
        try:
            lib.mra_qformer_destroy(self._handle)
        except Exception:
            pass

    def _params_version(self):
        ps = list(self.bert.parameters()) + list(self.llm_proj.parameters()) + [self.query_tokens]
        return tuple(p._version for p in ps)

    def sync_operands(self):
        """Refresh the bf16 operand copy if a parameter was written through PyTorch since the last refresh
        (``load_state_dict``, manual ``copy_``): in-place writes bump the parameter's version counter, whereas the
        optimizer's raw-pointer updates refresh the copy themselves."""
        v = self._params_version()
        if v != self._seen_params_version:
            self.refresh_operands()

    def weight_struct(self):
        """The C-ABI weight struct over the live training buffers (bf16 operand copy + fp32 vectors)."""
        return self._structs[0]

    def proj_operands(self):
        """(bf16 weight, fp32 bias) of llm_proj as the kernels read them: views of the live buffers."""
        return self._view(self.flat16, "w_proj"), self._view(self.flat, "b_proj")

    # ------------------------------------------------------------------------------------------------ flat views
    def _view(self, buf, name, row0=0, nrows=None):
        off, shp = self.seg[name]
        if nrows is None:
            return buf[off:off + math.prod(shp)].view(shp)
        inner = math.prod(shp[1:]) if len(shp) > 1 else 1
        return buf[off + row0 * inner: off + (row0 + nrows) * inner].view((nrows,) + tuple(shp[1:]))

    def _bind(self, param: nn.Parameter, name, row0=0, nrows=None):
        v = self._view(self.flat, name, row0, nrows)
        assert v.shape == param.shape, (name, v.shape, param.shape)
        v.copy_(param.data)
        param.data = v
        param.grad = self._view(self.grad, name, row0, nrows)

    def _ptr(self, buf, name, esize):
        return buf.data_ptr() + self.seg[name][0] * esize

    def refresh_operands(self):
        """bf16 operand copy of the master weights (one cast kernel; ``adam_step`` refreshes it in the same pass)."""
        check(lib.mra_cast_bf16(self.flat.data_ptr(), self.flat16.data_ptr(), self.numel, current_stream()))
        self.version += 1
        self._seen_params_version = self._params_version()
        if self._structs is None:
            self._structs = self._build_structs()
            check(lib.mra_qformer_set_weights(self._handle, C.byref(self._structs[0])))

    def _build_structs(self):
        W, G = _lib.QFormerWeights(), _lib.QFormerGrads()
        for f in ("word_emb", "pos_emb", "w_ckv", "w_proj"):
            setattr(W, f, self._ptr(self.flat16, f, 2))
        for f in ("ln_e_g", "ln_e_b", "b_ckv", "b_proj"):
            setattr(W, f, self._ptr(self.flat, f, 4))
        for f in ("word_emb", "pos_emb", "ln_e_g", "ln_e_b", "w_ckv", "b_ckv", "w_proj", "b_proj", "query_tokens"):
            setattr(G, f, self._ptr(self.grad, f, 4))
        for l, c in enumerate(self.cross):
            for n, shp in _layer_segments(self.H, self.I, self.W, c):
                key = f"L{l}.{n}"
                is_mat = len(shp) == 2
                setattr(W.layer[l], n, self._ptr(self.flat16, key, 2) if is_mat else self._ptr(self.flat, key, 4))
                setattr(G.layer[l], n, self._ptr(self.grad, key, 4))
        return W, G

    # ------------------------------------------------------------------------------------------------ forward / backward
    def forward(self, enc: torch.Tensor, input_ids: Optional[torch.Tensor], attention_mask: Optional[torch.Tensor],
                dropout_p: float = 0.0, dropout_seed: int = 0) -> torch.Tensor:
        """Training forward (activations kept): returns ``llm_proj(Qformer.bert(...).last_hidden_state[:, :Nq])`` as
        bf16 ``[rows, Nq, D]`` attached to autograd; ``.backward()`` accumulates into the flat gradient buffer.

        ``dropout_p > 0`` = ``model.train()`` of the reference (utils/trainer.py:110): Philox dropout on the embeddings, the
        attention probabilities and the attention-output / FFN-output Linears, mask regenerated (not stored) by the backward.
        The text is padded to a multiple of 32 tokens with masked-out padding (TMA attention kernels)."""
        self._dropout = (float(dropout_p), int(dropout_seed))
        return _QFormerTrainFn.apply(self, enc, input_ids, attention_mask, self.query_tokens)

    def _forward_impl(self, enc, input_ids, attention_mask):
        self.sync_operands()
        drop_p, drop_seed = getattr(self, "_dropout", (0.0, 0))
        self._dropout = (0.0, 0)
        if drop_p > 0.0 and input_ids is not None and input_ids.shape[1] % 32 != 0:
            pad = 32 - input_ids.shape[1] % 32
            Nq0 = self.cfg.query_length
            if attention_mask is None:
                attention_mask = torch.ones(input_ids.shape[0], Nq0 + input_ids.shape[1], dtype=torch.long, device=input_ids.device)
            input_ids = torch.nn.functional.pad(input_ids, (0, pad), value=0)
            attention_mask = torch.nn.functional.pad(attention_mask, (0, pad), value=0)
        cfg = self.cfg
        dev = enc.device
        rows, Nk, Wd = enc.shape
        if Wd != cfg.encoder_width:
            raise ValueError(f"encoder_hidden_states width {Wd} != config.encoder_width {cfg.encoder_width}")
        Nq = cfg.query_length
        enc_b = enc.to(torch.bfloat16).contiguous()
        T, ids, amask = 0, None, None
        if input_ids is not None:
            T = input_ids.shape[1]
            ids = input_ids.to(device=dev, dtype=torch.int32).contiguous()
        if attention_mask is not None:
            amask = attention_mask.to(device=dev, dtype=torch.int32).contiguous()
        flags = _lib.FWD_SAVE_FOR_BACKWARD | _lib.FWD_SKIP_DEAD_TEXT_FFN
        need = lib.mra_qformer_workspace_bytes(self._handle, rows, T, Nk, flags)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=dev, dtype=torch.uint8)
        out = torch.empty(rows * Nq, self.D, device=dev, dtype=torch.bfloat16)
        io = _lib.QFormerIO(enc=enc_b.data_ptr(), input_ids=_lib.ptr(ids), attn_mask=_lib.ptr(amask), enc_mask=None,
                            query_embeds=self._ptr(self.flat, "query_tokens", 4), q_rows=self.query_tokens.shape[0],
                            rows=rows, T=T, Nk=Nk, flags=flags, last_hidden=None, llm_out=out.data_ptr(),
                            dropout_p=drop_p, dropout_seed=drop_seed & 0xFFFFFFFFFFFFFFFF)
        check(lib.mra_qformer_forward(self._handle, C.byref(io), self._ws.data_ptr(), self._ws.numel(), current_stream()))
        self._saved = (io, enc_b, ids, amask, rows, T, Nk)
        return out.view(rows, Nq, self.D)

    def _backward_impl(self, d_out: torch.Tensor):
        io, enc_b, ids, amask, rows, T, Nk = self._saved
        d = d_out.reshape(rows * self.cfg.query_length, self.D).to(torch.bfloat16).contiguous()
        need = lib.mra_qformer_backward_workspace_bytes(self._handle, rows, T, Nk)
        if self._bws is None or self._bws.numel() < need:
            self._bws = torch.empty(need, device=d.device, dtype=torch.uint8)
        W, G = self._structs
        check(lib.mra_qformer_backward(self._handle, C.byref(io), d.data_ptr(), None, C.byref(G), self._ws.data_ptr(),
                                       self._ws.numel(), self._bws.data_ptr(), self._bws.numel(), current_stream()))
        self.last_backward_launches = lib.mra_qformer_last_launch_count(self._handle)
        self._saved = None
        self.backward_done.record(torch.cuda.current_stream())
        self.backward_ran = True

    def _exchange(self, lo: int, hi: int):
        """all-reduce (sum) of one range of the flat gradient buffer on the current (communication) stream"""
        import torch.distributed as dist
        if self.grad_comm_dtype == torch.bfloat16:
            if self.grad16 is None:
                self.grad16 = torch.empty(self.numel, device=self.grad.device, dtype=torch.bfloat16)
            check(lib.mra_cast_bf16(self.grad.data_ptr() + lo * 4, self.grad16.data_ptr() + lo * 2, hi - lo, current_stream()))
            dist.all_reduce(self.grad16[lo:hi], op=dist.ReduceOp.SUM, group=self.reduce_group)
        else:
            dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.reduce_group)

    def exchange_bucket(self, k: int, comm_stream: torch.cuda.Stream):
        """Enqueue the all-reduce of bucket ``k`` on ``comm_stream`` behind the event that makes its gradients final."""
        ev, ranges = self.buckets[k]
        comm_stream.wait_event(self.layer_events[ev] if ev is not None else self.backward_done)
        with torch.cuda.stream(comm_stream):
            for lo, hi in ranges:
                self._exchange(lo, hi)
        self._reduced_bf16 = self.grad_comm_dtype == torch.bfloat16

    def flat_grad_allreduce(self):
        """one exchange of the whole buffer on the current stream, after the backward (the A/B baseline of the overlap)"""
        self._exchange(0, self.numel)
        self._reduced_bf16 = self.grad_comm_dtype == torch.bfloat16

    # ------------------------------------------------------------------------------------------------ optimizer
    def zero_grad(self):
        self.grad.zero_()

    def adam_step(self, lr: float, grad_scale: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                  zero_grad: bool = False):
        """``optimizer.step()`` (+ ``optimizer.zero_grad()`` with ``zero_grad``) and the refresh of the bf16 operand copy
        in ONE pass over the flat buffers."""
        self.step_count += 1
        self.version += 1
        g16 = self.grad16.data_ptr() if self._reduced_bf16 else None      # gradients as exchanged in bf16 (else the fp32 buffer)
        self._reduced_bf16 = False
        check(lib.mra_adam_step_fused(self.flat.data_ptr(), self.grad.data_ptr(), g16, self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.flat16.data_ptr(), self.numel, lr, betas[0], betas[1], eps,
                                      weight_decay, self.step_count, grad_scale, int(zero_grad), current_stream()))

    def begin_bucketed_step(self):
        """Start an optimizer step that is applied bucket by bucket (``adam_bucket``): one step count for all buckets."""
        self.step_count += 1
        self.version += 1

    def adam_bucket(self, k: int, lr: float, grad_scale: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8,
                    weight_decay: float = 0.0, zero_grad: bool = True, reduced_bf16: bool = False, hyper_dev: Optional[int] = None):
        """``adam_step`` restricted to the flat-buffer ranges of gradient bucket ``k`` (same arithmetic per element), on the
        current stream.  Called behind the event that makes the bucket's gradients final (and behind its exchange), so the
        update of the top layers runs under the backward of the lower ones: the backward's launches at fine-tuning batch
        sizes leave most SMs and nearly all of the HBM bandwidth idle, which is what this 28-bytes-per-parameter pass needs.
        The bf16 operand copy of a bucket is rewritten here too: no kernel of this step reads those weights any more."""
        for lo, hi in self.buckets[k][1]:
            g16 = self.grad16.data_ptr() + lo * 2 if reduced_bf16 else None
            if hyper_dev is not None:   # lr / bias corrections / gradient scale from device memory (CUDA-graph replay)
                check(lib.mra_adam_step_fused_dyn(self.flat.data_ptr() + lo * 4, self.grad.data_ptr() + lo * 4, g16,
                                                  self.exp_avg.data_ptr() + lo * 4, self.exp_avg_sq.data_ptr() + lo * 4,
                                                  self.flat16.data_ptr() + lo * 2, hi - lo, betas[0], betas[1], eps, weight_decay,
                                                  int(zero_grad), hyper_dev, current_stream()))
                continue
            check(lib.mra_adam_step_fused(self.flat.data_ptr() + lo * 4, self.grad.data_ptr() + lo * 4, g16,
                                          self.exp_avg.data_ptr() + lo * 4, self.exp_avg_sq.data_ptr() + lo * 4,
                                          self.flat16.data_ptr() + lo * 2, hi - lo, lr, betas[0], betas[1], eps, weight_decay,
                                          self.step_count, grad_scale, int(zero_grad), current_stream()))


class _QFormerTrainFn(torch.autograd.Function):
    """Autograd node of one modality's Q-Former + projection.  The parameter gradients are accumulated by the CUDA
    backward directly into ``TrainableQFormer.grad`` (which the parameters' ``.grad`` alias), so ``None`` is returned for
    them; ``query_tokens`` is an input only to make autograd schedule the node."""

    @staticmethod
    def forward(ctx, state: TrainableQFormer, enc, input_ids, attention_mask, query_tokens):
        ctx.state = state
        return state._forward_impl(enc, input_ids, attention_mask)

    @staticmethod
    def backward(ctx, d_out):
        ctx.state._backward_impl(d_out)
        return None, None, None, None, None


def dist_rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def warmup_cosine_lr(cur_epoch: int, cur_step: int, max_epoch: int, init_lr: float = 3e-4, min_lr: float = 0.0,
                     warmup_steps: int = 1000, warmup_start_lr: float = 1e-8) -> float:
    """LAVIS ``LinearWarmupCosineLRScheduler.step`` (utils/trainer.py:66,127): linear warm-up over the first
    ``warmup_steps`` iterations of epoch 0, then a per-epoch cosine."""
    if cur_epoch == 0 and cur_step < warmup_steps:
        return min(init_lr, warmup_start_lr + (init_lr - warmup_start_lr) * cur_step / max(warmup_steps, 1))
    return (init_lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * cur_epoch / max_epoch)) + min_lr


class _StepGraphs:
    """The captured CUDA graphs of one input shape of ``QFormerTrainer.train_step``: ``fwd`` (input preparation + the
    training forward of every modality on its stream) and, per (stepping, reduce) variant, the backward of every modality
    on its stream + the bucketed gradient exchange + the per-bucket Adam update on the communication stream.  Inputs,
    projected outputs and their gradients live in static buffers; the optimizer's per-step scalars in device memory."""

    def __init__(self, tr: "QFormerTrainer", mods, feats, input_ids, attention_mask):
        self.tr, self.mods = tr, mods
        dev = next(tr.model.parameters()).device
        self.feats = {m: torch.empty(feats[m].shape, dtype=feats[m].dtype, device=dev) for m in mods}
        self.ids = torch.empty(input_ids.shape, dtype=input_ids.dtype, device=dev)
        self.mask = torch.empty(attention_mask.shape, dtype=attention_mask.dtype, device=dev)
        for m in mods:
            self.feats[m].copy_(feats[m])
        self.ids.copy_(input_ids)
        self.mask.copy_(attention_mask)
        self.hyper_host = {m: torch.zeros(4, dtype=torch.float32) for m in mods}   # pageable: the H2D copy below stages it before returning
        self.hyper_dev = {m: torch.zeros(4, dtype=torch.float32, device=dev) for m in mods}
        self.bwd: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.keep = []
        # eager warm-up of exactly what is captured (lazy initialisation: kernel attributes, workspaces, tensor maps): the
        # backward runs on a zero gradient, i.e. accumulates nothing, and no optimizer step is taken
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            y = self._forward()
            for m in mods:
                tr.states[m]._backward_impl(torch.zeros(y[m].shape, dtype=torch.bfloat16, device=dev))
            self._join()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        self.fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd, capture_error_mode="thread_local"):
            self.y = self._forward()
            self._join()
        self.saved = {m: tr.states[m]._saved for m in mods}
        self.ws = [(tr.states[m]._ws, tr.states[m]._bws) for m in mods]      # the graphs hold raw pointers into these
        Nq = tr.model.num_query_token
        self.dy = {m: torch.zeros(self.y[m].shape[0] * self.y[m].shape[1], self.y[m].shape[2], dtype=torch.bfloat16, device=dev)
                   for m in mods}

    def _forward(self):
        """``QFormerTrainer.forward_modalities`` on the static inputs (no autograd node, no record_stream)."""
        tr, model = self.tr, self.tr.model
        main = torch.cuda.current_stream()
        ys, self._sides = {}, []
        for m in self.mods:
            enc = model.fold_frames(m, self.feats[m], apply_ln=False)
            bs = self.ids.shape[0]
            num = enc.shape[0] // bs
            ids = self.ids.repeat(num, 1)
            tmask = self.mask.repeat(num, 1)
            q_atts = torch.ones(enc.shape[0], model.num_query_token, dtype=tmask.dtype, device=tmask.device)
            full_mask = torch.cat([q_atts, tmask], 1)
            st = tr.states[m]
            st._dropout = (0.0, 0)
            if tr.parallel_modalities and len(self.mods) > 1:
                side = tr.mod_streams[m]
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    y = st._forward_impl(enc, ids, full_mask)
                self._sides.append(side)
            else:
                y = st._forward_impl(enc, ids, full_mask)
            self.keep += [enc, ids, full_mask]
            ys[m] = y.reshape(bs, num * model.num_query_token, -1)
        return ys

    def _join(self):
        for side in self._sides:
            torch.cuda.current_stream().wait_stream(side)

    def _capture_backward(self, stepping: bool, reduce: bool, world: int) -> torch.cuda.CUDAGraph:
        tr = self.tr
        assert not reduce, "the gradient exchange is not captured (single-process replay only)"
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.fwd.pool(), capture_error_mode="thread_local"):
            main = torch.cuda.current_stream()
            sides = []
            for m in self.mods:
                st = tr.states[m]
                st._saved = self.saved[m]
                if tr.parallel_modalities and len(self.mods) > 1:
                    side = tr.mod_streams[m]
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        st._backward_impl(self.dy[m])
                    sides.append(side)
                else:
                    st._backward_impl(self.dy[m])
            live = [tr.states[m] for m in self.mods]
            if stepping:
                for k in range(max(len(st.buckets) for st in live)):
                    for m, st in zip(self.mods, live):
                        if k >= len(st.buckets):
                            continue
                        ev = st.buckets[k][0]
                        tr.comm_stream.wait_event(st.layer_events[ev] if ev is not None else st.backward_done)
                        with torch.cuda.stream(tr.comm_stream):
                            st.adam_bucket(k, 0.0, zero_grad=True, hyper_dev=self.hyper_dev[m].data_ptr())
                main.wait_stream(tr.comm_stream)
            for side in sides:
                main.wait_stream(side)
        return g

    def replay_backward(self, stepping: bool, reduce: bool, world: int):
        tr = self.tr
        key = (stepping, reduce, tuple(tr.states[m].grad_comm_dtype for m in self.mods))
        g = self.bwd.get(key)
        if g is None:
            g = self.bwd[key] = self._capture_backward(stepping, reduce, world)
        if stepping:
            for m in self.mods:
                st = tr.states[m]
                st.begin_bucketed_step()
                lib.mra_adam_hyper(tr.lr, 0.9, 0.999, st.step_count, 1.0 / world,
                                   C.cast(self.hyper_host[m].data_ptr(), C.POINTER(C.c_float)))
                self.hyper_dev[m].copy_(self.hyper_host[m])
        g.replay()


class QFormerTrainer:
    """The hot-path slice of ``utils/trainer.py::Trainer`` for ``XInstructBLIPQFormers``: ``train_step`` = one iteration of
    ``train_epoch`` (:124-140).  ``loss_fn(inputs_llm, atts_llm, samples) -> scalar`` stands in for the frozen LLM's loss
    (out of scope); the default is the surrogate ``sum(inputs_llm * G)`` used by the parity tests and the benchmark."""

    def __init__(self, model, max_epoch: int = 1, accum_grad_iters: int = 2, init_lr: float = 3e-4, warmup_steps: int = 1000,
                 loss_fn=None, group=None, overlap_allreduce: bool = True, parallel_modalities: bool = True,
                 grad_comm_dtype: torch.dtype = torch.bfloat16, dropout: Optional[float] = 0.0, seed: int = 0,
                 cuda_graph: bool = False):
        self.model = model
        model.freeze_qformers(False)
        self.states = {m: TrainableQFormer(getattr(model, f"{m}_Qformer"), getattr(model, f"{m}_query_tokens"),
                                           getattr(model, f"{m}_llm_proj")) for m in model.modalities}
        self.accum_grad_iters, self.max_epoch, self.init_lr, self.warmup_steps = accum_grad_iters, max_epoch, init_lr, warmup_steps
        self.loss_fn = loss_fn
        self.group = group
        self.overlap_allreduce = overlap_allreduce
        # dropout: 0.0 (default: the parity configuration) or a probability; None = the Q-Former config's
        # hidden_dropout_prob (0.1), i.e. the reference's model.train() (utils/trainer.py:110)
        self.dropout = dropout
        self.seed = seed
        self.allreduce_enabled = True       # False: skip the gradient exchange (bench.py measures its exposed cost that way)
        # Adam per gradient bucket behind the backward, on the communication stream (TrainableQFormer.adam_bucket), instead of
        # one pass after it.  Off by default: measured neutral on B200 (13.77 vs 13.77 ms at 1 GPU, 14.95 vs 14.77 ms at 2 --
        # the backward's launches do not leave the HBM bandwidth idle that the update needs); the CUDA-graph replay uses it.
        self.overlap_optimizer = False
        self.set_grad_comm_dtype(grad_comm_dtype)
        # At fine-tuning batch sizes (64 rows per modality) most launches fill a fraction of the 148 SMs, so the video and
        # the audio Q-Former run on two streams: autograd replays each modality's backward on the stream of its forward.
        # (For the large-batch inference forward the opposite holds -- DESIGN.md section 4 -- and one stream is used.)
        self.parallel_modalities = parallel_modalities
        dev = next(model.parameters()).device
        self.mod_streams = {m: torch.cuda.Stream(dev) for m in model.modalities}
        self.comm_stream = torch.cuda.Stream(dev)     # ONE stream for the gradient exchange of all modalities
        self.iter = 0
        self.lr = init_lr
        # cuda_graph: replay the step from two captured CUDA graphs (forward | backward + gradient exchange + optimizer)
        # around the eagerly evaluated loss.  The ~800 launches of a config-4 step take the host ~10 ms to issue against
        # ~14 ms on the device, so the streams of the two modalities starve each other; a replay costs the host two
        # launches.  One pair of graphs per input shape (batches of a fixed shape: pad the text to a fixed length); needs
        # dropout 0 and falls back to the eager path otherwise.
        self.cuda_graph = cuda_graph
        self._graphs: Dict[tuple, "_StepGraphs"] = {}
        self.max_graphs = 4

    def set_grad_comm_dtype(self, dtype: torch.dtype):
        """bf16 (default: 0.74 GB per step for both Q-Formers) or fp32 (1.49 GB, exact) gradient exchange"""
        assert dtype in (torch.bfloat16, torch.float32)
        for st in self.states.values():
            st.grad_comm_dtype = dtype

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def forward_modalities(self, feats, input_ids, attention_mask):
        """Training-mode ``encode_modalities`` (models/xinstructblip.py:456-477) -> (inputs_llm, atts_llm)."""
        model = self.model
        inputs_llm, atts_llm = {}, {}
        joins = []
        for m in model.modalities:
            if m not in feats:
                continue
            enc = model.fold_frames(m, feats[m], apply_ln=False)
            bs = input_ids.shape[0]
            num = enc.shape[0] // bs
            ids = input_ids.repeat(num, 1)                      # reference tiling (:463)
            tmask = attention_mask.repeat(num, 1)
            q_atts = torch.ones(enc.shape[0], model.num_query_token, dtype=tmask.dtype, device=tmask.device)
            full_mask = torch.cat([q_atts, tmask], 1)
            p_drop = getattr(model, f"{m}_Qformer").config.hidden_dropout_prob if self.dropout is None else self.dropout
            # one dropout stream per (trainer seed, iteration, modality, rank): ranks and modalities draw independent masks
            dkw = {}
            if p_drop and p_drop > 0.0:
                rank = dist_rank() if self._world() > 1 else 0
                dkw = dict(dropout_p=p_drop, dropout_seed=((self.seed * 1000003 + self.iter) * 64 + model.modalities.index(m)) * 4096 + rank)
            if self.parallel_modalities and len(model.modalities) > 1:
                main, side = torch.cuda.current_stream(), self.mod_streams[m]
                side.wait_stream(main)                      # inputs prepared on the main stream
                with torch.cuda.stream(side):
                    y = self.states[m].forward(enc, ids, full_mask, **dkw)
                for t in (enc, ids, full_mask):
                    t.record_stream(side)
                joins.append(side)
            else:
                y = self.states[m].forward(enc, ids, full_mask, **dkw)
            inputs_llm[m] = y.reshape(bs, num, model.num_query_token, -1).view(bs, num * model.num_query_token, -1)
            atts_llm[m] = torch.ones(inputs_llm[m].size()[:-1], dtype=torch.long, device=y.device)
        for side in joins:                                  # the loss is computed on the main stream
            torch.cuda.current_stream().wait_stream(side)
        return inputs_llm, atts_llm

    def train_step(self, feats, input_ids, attention_mask, samples=None, cur_epoch: int = 0, surrogate: Optional[Dict[str, torch.Tensor]] = None,
                   apply_optimizer: bool = True):
        """One iteration of the reference's inner loop.  Returns the (unscaled) loss tensor.  ``apply_optimizer=False`` stops
        after the gradient all-reduce of a stepping iteration (the flat gradient buffers then hold the SUM over ranks):
        used by the data-parallel equivalence checks."""
        import torch.distributed as dist
        self.lr = warmup_cosine_lr(cur_epoch, self.iter, self.max_epoch, self.init_lr, 0.0, self.warmup_steps)   # :127
        if self.cuda_graph and apply_optimizer and self._graph_eligible(feats):
            return self._train_step_graphed(feats, input_ids, attention_mask, samples, surrogate)
        inputs_llm, atts_llm = self.forward_modalities(feats, input_ids, attention_mask)
        if self.loss_fn is not None:
            loss = self.loss_fn(inputs_llm, atts_llm, samples)
        else:
            loss = sum((inputs_llm[m].float() * surrogate[m]).sum() for m in inputs_llm)
        world = self._world()
        stepping = (self.iter + 1) % self.accum_grad_iters == 0                                                    # :137
        reduce = stepping and world > 1 and self.allreduce_enabled
        # optimizer applied bucket by bucket behind the backward (see TrainableQFormer.adam_bucket)
        bucketed_opt = stepping and apply_optimizer and self.overlap_allreduce and self.overlap_optimizer
        for st in self.states.values():
            st.backward_ran = False
        (loss / self.accum_grad_iters).backward()                                                                  # :131-133
        # (the CUDA backward of each modality ran on its own stream and accumulated into the flat gradient buffers behind
        #  autograd's back -- no AccumulateGrad node: the streams are joined explicitly below)
        live = [st for st in self.states.values() if st.backward_ran]
        bucketed = set()
        if (reduce or bucketed_opt) and self.overlap_allreduce and live:
            # DDP's gradient averaging (utils/trainer.py:69), only on optimizer-step iterations (the reference all-reduces on
            # every backward).  By now the whole backward of every modality is ENQUEUED (the host runs far ahead of the GPU)
            # and has recorded, per bucket, the event that makes its gradients final.  The buckets of all modalities go on
            # one communication stream in the order in which they become ready -- projection, top layers, ..., tail of
            # modality A, tail of modality B alternating -- so that no modality's exchange queues behind another's tail;
            # each bucket's Adam update follows its exchange on the same stream, under the rest of the backward.
            # Sum semantics; grad_scale = 1 / world turns it into DDP's mean.
            for st in live:
                st.reduce_group = self.group
                if bucketed_opt:
                    st.begin_bucketed_step()
                    bucketed.add(id(st))
            for k in range(max(len(st.buckets) for st in live)):
                for st in live:
                    if k >= len(st.buckets):
                        continue
                    if reduce:
                        st.exchange_bucket(k, self.comm_stream)
                    else:
                        ev = st.buckets[k][0]
                        self.comm_stream.wait_event(st.layer_events[ev] if ev is not None else st.backward_done)
                    if bucketed_opt:
                        with torch.cuda.stream(self.comm_stream):
                            st.adam_bucket(k, self.lr, grad_scale=1.0 / world, zero_grad=True,
                                           reduced_bf16=reduce and st.grad_comm_dtype == torch.bfloat16)
            if bucketed_opt:
                for st in live:
                    st._reduced_bf16 = False
        if self.parallel_modalities and len(self.mod_streams) > 1:
            for side in self.mod_streams.values():
                torch.cuda.current_stream().wait_stream(side)
        self.iter += 1
        if stepping:
            if (reduce and self.overlap_allreduce) or bucketed:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            if reduce:
                for st in self.states.values():
                    st.reduce_group = self.group
                    if not (self.overlap_allreduce and st.backward_ran):
                        # flat all-reduce after the backward: the A/B baseline, and modalities absent from this step (their
                        # accumulated gradients of earlier iterations still have to be averaged)
                        st.flat_grad_allreduce()
                if not apply_optimizer:
                    for st in self.states.values():       # leave the exchanged SUM in the fp32 buffers for the caller
                        if st._reduced_bf16:
                            st.grad.copy_(st.grad16)
                            st._reduced_bf16 = False
            if apply_optimizer:
                for st in self.states.values():
                    if id(st) not in bucketed:
                        st.adam_step(self.lr, grad_scale=1.0 / world, zero_grad=True)
        return loss.detach()

    # ------------------------------------------------------------------------------------------------ CUDA-graph replay
    def _graph_eligible(self, feats) -> bool:
        for m in self.model.modalities:
            if m in feats:
                p = getattr(self.model, f"{m}_Qformer").config.hidden_dropout_prob if self.dropout is None else self.dropout
                if p and p > 0.0:
                    return False      # the dropout seed of an iteration is a launch argument: not replayable
        if self._world() > 1:
            return False          # single-process only: a captured NCCL exchange hung at 2 ranks (profiles/r02_NOTES.md);
                                  # the eager step keeps the exchange overlapped with the backward
        return self.overlap_allreduce and all(t.is_cuda or t.is_pinned() for t in feats.values())

    def _train_step_graphed(self, feats, input_ids, attention_mask, samples, surrogate):
        world = self._world()
        stepping = (self.iter + 1) % self.accum_grad_iters == 0
        reduce = stepping and world > 1 and self.allreduce_enabled
        mods = tuple(m for m in self.model.modalities if m in feats)
        key = (mods, tuple((tuple(feats[m].shape), feats[m].dtype) for m in mods), tuple(input_ids.shape),
               tuple(attention_mask.shape), self.parallel_modalities)
        G = self._graphs.get(key)
        if G is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            G = self._graphs[key] = _StepGraphs(self, mods, feats, input_ids, attention_mask)
        for m in mods:
            if feats[m].data_ptr() != G.feats[m].data_ptr():
                G.feats[m].copy_(feats[m], non_blocking=True)
            self.states[m].sync_operands()
        G.ids.copy_(input_ids, non_blocking=True)
        G.mask.copy_(attention_mask, non_blocking=True)
        G.fwd.replay()
        # the loss stays eager (any user function of the projected query tokens: the frozen LLM in the reference)
        leaf = {m: G.y[m].detach().requires_grad_(True) for m in mods}
        atts = {m: torch.ones(leaf[m].size()[:-1], dtype=torch.long, device=leaf[m].device) for m in mods}
        if self.loss_fn is not None:
            loss = self.loss_fn(leaf, atts, samples)
        else:
            loss = sum((leaf[m].float() * surrogate[m]).sum() for m in mods)
        grads = torch.autograd.grad(loss / self.accum_grad_iters, [leaf[m] for m in mods])
        for m, g in zip(mods, grads):
            G.dy[m].copy_(g.reshape(G.dy[m].shape))
        G.replay_backward(stepping, reduce, world)
        self.iter += 1
        return loss.detach()

    def eval_epoch(self, generations, group=None):
        """``Trainer.eval_epoch`` (utils/trainer.py:156-182) from the point where the LLM has produced its strings:
        ``generations`` is this rank's iterable of ``(qid, query, vid, target_text, output_text)``; targets and outputs go
        through ``moment_str_to_list(post_process(.))`` (:168-169), the records are scored on this rank's GPU, and -- unlike
        the reference, which evaluates only the local shard on every rank -- gathered to rank 0 (one fixed-width gather,
        ``mr_eval.score_records_distributed``) so that rank 0 returns the metrics of the WHOLE validation set
        (``eval_submission(results, results)``, :181); other ranks return None."""
        import torch.distributed as dist
        from . import mr_eval
        from .parsing import parse_output
        group = group if group is not None else self.group
        results = []
        for qid, query, vid, target, output in generations:
            results.append({"qid": qid, "query": query, "vid": vid, "relevant_windows": parse_output(target),
                            "pred_relevant_windows": parse_output(output)})
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if not distributed:
            return mr_eval.eval_submission(results, results, verbose=False) if results else None
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        # global submission order = (rank, local index); qids are made unique per rank for the rank-0 bookkeeping
        counts = [None] * world
        dist.all_gather_object(counts, len(results), group=group)
        base = sum(counts[:rank])
        mine = [dict(d, _order=base + i) for i, d in enumerate(results)]
        rec = mr_eval.score_records_distributed(mine, mine, group=group)
        if rank != 0:
            return None
        total = sum(counts)
        stub = [{"qid": i, "pred_relevant_windows": [[0, 0]], "relevant_windows": [[0, 0]]} for i in range(total)]
        return mr_eval.eval_submission(stub, stub, verbose=False, _records=rec)

    def state_dict_trainable(self):
        """``_save_checkpoint`` payload (utils/trainer.py:184-199): only parameters that require grad."""
        grad = {k: v.requires_grad for k, v in self.model.named_parameters()}
        return {k: v for k, v in self.model.state_dict().items() if grad.get(k, False)}

    # ------------------------------------------------------------------------------------------------ checkpoints
    # ``checkpoint["optimizer"]`` has the layout of ``torch.optim.Adam(model.parameters(), lr=3e-4).state_dict()`` -- what
    # the reference saves and loads (utils/trainer.py:65,199,255): ``state[i] = {step, exp_avg, exp_avg_sq}`` for the i-th
    # parameter of ``model.parameters()`` and one ``param_groups`` entry -- so a stock Adam over this module can resume from
    # it and vice versa.  The moments are views of / copied into the flat buffers; the (shared) step count is per modality.
    def _param_slices(self):
        """[(index in model.parameters(), name, state, offset, numel)] for every parameter that lives in a flat buffer."""
        out = []
        for i, (name, p) in enumerate(self.model.named_parameters()):
            for st in self.states.values():
                off = (p.data_ptr() - st.flat.data_ptr()) // 4
                if p.device == st.flat.device and 0 <= off < st.numel and p.data_ptr() >= st.flat.data_ptr():
                    out.append((i, name, st, int(off), p.numel()))
                    break
        return out

    def optimizer_state_dict(self, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        n_params = sum(1 for _ in self.model.parameters())
        state = {}
        for i, name, st, off, n in self._param_slices():
            state[i] = {"step": torch.tensor(float(st.step_count)),
                        "exp_avg": st.exp_avg[off:off + n].clone(), "exp_avg_sq": st.exp_avg_sq[off:off + n].clone()}
        shapes = {i: p.shape for i, p in enumerate(self.model.parameters())}
        for i, d in state.items():
            d["exp_avg"] = d["exp_avg"].view(shapes[i])
            d["exp_avg_sq"] = d["exp_avg_sq"].view(shapes[i])
        if all(st.step_count == 0 for st in self.states.values()):
            state = {}   # torch's Adam has no per-parameter state before its first step
        group = {"lr": self.lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(n_params))}
        return {"state": state, "param_groups": [group],
                "param_names": [n for n, _ in self.model.named_parameters()]}   # extra key: makes the indices self-describing

    def load_optimizer_state_dict(self, osd):
        if not isinstance(osd, dict) or "state" not in osd or "param_groups" not in osd:
            raise ValueError("checkpoint['optimizer'] is not a torch.optim.Adam state_dict (keys 'state', 'param_groups'); "
                             "round-1 checkpoints of this package ({modality: {exp_avg, exp_avg_sq, step}}) must be re-saved")
        names = [n for n, _ in self.model.named_parameters()]
        ck_names = osd.get("param_names")
        n_ck = sum(len(g["params"]) for g in osd["param_groups"])
        if ck_names is None and n_ck != len(names):
            raise ValueError(f"optimizer state covers {n_ck} parameters but this module has {len(names)}: it was saved over a "
                             "different parameter list (e.g. the reference's full model incl. the LLM) and carries no "
                             "'param_names' to map it; resume the Q-Former weights from checkpoint['model'] only")
        index_of = {n: i for i, n in enumerate(ck_names)} if ck_names is not None else {n: i for i, n in enumerate(names)}
        steps = {id(st): [] for st in self.states.values()}
        for _, name, st, off, n in self._param_slices():
            d = osd["state"].get(index_of.get(name, -1))
            if d is None:
                st.exp_avg[off:off + n].zero_()
                st.exp_avg_sq[off:off + n].zero_()
                continue
            if d["exp_avg"].numel() != n:
                raise ValueError(f"optimizer state of {name}: {tuple(d['exp_avg'].shape)} does not match the parameter")
            st.exp_avg[off:off + n].copy_(d["exp_avg"].reshape(-1))
            st.exp_avg_sq[off:off + n].copy_(d["exp_avg_sq"].reshape(-1))
            steps[id(st)].append(int(float(d["step"])))
        for st in self.states.values():
            got = steps[id(st)]
            st.step_count = max(got) if got else 0
        if osd["param_groups"]:
            self.lr = float(osd["param_groups"][0].get("lr", self.lr))

    def load_checkpoint(self, path: str) -> int:
        """``Trainer._load_checkpoint`` (utils/trainer.py:236-260): parameters (only the trainable ones were saved), the
        Adam state and the epoch; returns the epoch to resume from (``checkpoint["epoch"] + 1``)."""
        ckpt = torch.load(path, map_location=next(self.model.parameters()).device)
        with torch.no_grad():
            sd = self.model.state_dict()
            for k, v in ckpt["model"].items():
                if k in sd:
                    sd[k].copy_(v)          # in place: the parameters are views of the flat master buffers
        self.load_optimizer_state_dict(ckpt["optimizer"])
        for st in self.states.values():
            st.refresh_operands()
        return int(ckpt["epoch"]) + 1

    def save_checkpoint(self, path: str, cur_epoch: int):
        """``Trainer._save_checkpoint`` (utils/trainer.py:184-210): {"model": trainable parameters only, "optimizer":
        Adam state_dict, "scaler": None (bf16 needs no loss scaling), "epoch"}."""
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"model": self.state_dict_trainable(), "optimizer": self.optimizer_state_dict(), "scaler": None,
                    "epoch": cur_epoch}, path)
