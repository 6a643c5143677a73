"""LLM-prompt assembly: the step right after the Q-Former path.

Reference: ``models/xinstructblip.py:342-385`` (generate) and ``:544-594`` (forward).  Per video the LLM input is

    for pos in range(F):  [enumeration tokens (a), (b) ...]  cue_video  32 video tokens  cue_audio  32 audio tokens
                          [timestamp tokens]
    duration tokens   prompt (+ answer) tokens

built there from ``4*F + 2`` ``torch.cat`` inputs and per-frame ``repeat``s, i.e. one more full copy of the largest
tensors of the path (the projected query tokens).  Here the layout is computed once (``PromptLayout``), ``llm_proj``'s
GEMM epilogue stores each frame's 32 tokens straight into its slot of ``inputs_embeds [bs, L, D]`` through a 4-D TMA
tensor map (``mra_qformer_io::llm_frames``), and every other piece -- embeddings of LLM tokens, produced by the caller's
frozen LLM embedding table, which is outside this path -- is copied to its slot by ONE ``mra_prompt_assemble`` launch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, lib

MODALITY_ORDER = ("video", "audio")   # the joint loop hard-codes this order (models/xinstructblip.py:359, :560)


@dataclass
class PromptPieces:
    """Embeddings / masks of the LLM-token pieces of the prompt, as the reference computes them with
    ``llm_model.get_input_embeddings()(tokens.input_ids)`` and ``tokens.attention_mask`` (all bf16 / int64, on the GPU).

    cue_embeds[m]        ``[Lc_m, D]``      ``emb_cue[m]`` (:366, same for every video and frame)
    cue_atts[m]          ``[Lc_m]``         ``tokenized_cue[m].attention_mask``
    timestamp_embeds     ``[bs, F, Tt, D]`` (:336-338) or None when ``interleave_seconds`` is off
    timestamp_atts       ``[bs, F, Tt]``
    enumeration_embeds   list over pos of ``[Te_pos, D]`` (:349-357; same for every video) or None
    duration_embeds      ``[bs, Td, D]``, duration_atts ``[bs, Td]`` (:372-381)
    prompt_embeds        ``[bs, Tp, D]``, prompt_atts ``[bs, Tp]`` (:384-386; prompt + answer in ``forward``)
    """
    cue_embeds: Dict[str, torch.Tensor]
    cue_atts: Dict[str, torch.Tensor]
    duration_embeds: torch.Tensor
    duration_atts: torch.Tensor
    prompt_embeds: torch.Tensor
    prompt_atts: torch.Tensor
    timestamp_embeds: Optional[torch.Tensor] = None
    timestamp_atts: Optional[torch.Tensor] = None
    enumeration_embeds: Optional[Sequence[torch.Tensor]] = None


@dataclass
class PromptLayout:
    """Row offsets of every piece inside one video's ``[L, D]`` block."""
    bs: int
    frames: int
    num_query: int
    D: int
    modalities: Tuple[str, ...]
    frame_start: List[int] = field(default_factory=list)          # first row of frame block f
    enum_len: List[int] = field(default_factory=list)
    cue_off: Dict[str, int] = field(default_factory=dict)         # offsets relative to (frame_start + enum_len)
    query_off: Dict[str, int] = field(default_factory=dict)
    ts_off: int = 0
    ts_len: int = 0
    body_len: int = 0                                             # frame block length without the enumeration tokens
    duration_start: int = 0
    prompt_start: int = 0
    L: int = 0

    @classmethod
    def build(cls, pieces: PromptPieces, bs: int, frames: int, num_query: int, modalities: Sequence[str]) -> "PromptLayout":
        D = pieces.prompt_embeds.shape[-1]
        mods = tuple(m for m in MODALITY_ORDER if m in modalities)
        lay = cls(bs=bs, frames=frames, num_query=num_query, D=D, modalities=mods)
        off = 0
        for m in mods:
            lay.cue_off[m] = off
            off += pieces.cue_embeds[m].shape[0]
            lay.query_off[m] = off
            off += num_query
        lay.ts_off = off
        lay.ts_len = pieces.timestamp_embeds.shape[2] if pieces.timestamp_embeds is not None else 0
        lay.body_len = off + lay.ts_len
        if pieces.enumeration_embeds is not None:
            if len(pieces.enumeration_embeds) != frames:
                raise ValueError(f"{len(pieces.enumeration_embeds)} enumeration pieces for {frames} frames")
            lay.enum_len = [int(e.shape[0]) for e in pieces.enumeration_embeds]
        else:
            lay.enum_len = [0] * frames
        row = 0
        for f in range(frames):
            lay.frame_start.append(row)
            row += lay.enum_len[f] + lay.body_len
        lay.duration_start = row
        lay.prompt_start = row + pieces.duration_embeds.shape[1]
        lay.L = lay.prompt_start + pieces.prompt_embeds.shape[1]
        return lay

    def body_start(self, f: int) -> int:
        return self.frame_start[f] + self.enum_len[f]

    @property
    def uniform(self) -> bool:
        """True when the frame bodies are equally spaced (always, unless enumeration pieces of pos >= 1 differ in length):
        then one strided descriptor addresses a piece in all frames."""
        if self.frames <= 2:
            return True
        d = self.body_start(1) - self.body_start(0)
        return all(self.body_start(f + 1) - self.body_start(f) == d for f in range(self.frames - 1))

    @property
    def frame_stride(self) -> int:
        return self.body_start(1) - self.body_start(0) if self.frames > 1 else self.body_len


class _Seg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_video_stride", C.c_int64), ("src_frame_stride", C.c_int64), ("rows", C.c_int32),
                ("frames", C.c_int32), ("dst_row", C.c_int32), ("dst_frame_rows", C.c_int32)]


def _segments(pieces: PromptPieces, lay: PromptLayout, dense_llm: Optional[Dict[str, torch.Tensor]]):
    """(src tensor, video stride, frame stride, rows, frames, dst_row, dst_frame_rows) of every copied piece."""
    D, F = lay.D, lay.frames
    segs = []

    def per_frame(src, vs, fs, rows, rel):
        if rows == 0:
            return
        if lay.uniform:
            segs.append((src, vs, fs, rows, F, lay.body_start(0) + rel, lay.frame_stride))
        else:
            for f in range(F):
                segs.append((src[:, f] if fs else src, vs, 0, rows, 1, lay.body_start(f) + rel, 0))

    for m in lay.modalities:
        per_frame(pieces.cue_embeds[m], 0, 0, pieces.cue_embeds[m].shape[0], lay.cue_off[m])
        if dense_llm is not None:   # (only without the scatter epilogue)
            y = dense_llm[m]        # [bs, F, Nq, D]
            per_frame(y, F * lay.num_query * D, lay.num_query * D, lay.num_query, lay.query_off[m])
    if lay.ts_len:
        t = pieces.timestamp_embeds
        per_frame(t, F * lay.ts_len * D, lay.ts_len * D, lay.ts_len, lay.ts_off)
    if pieces.enumeration_embeds is not None:
        for f, e in enumerate(pieces.enumeration_embeds):
            if e.shape[0]:
                segs.append((e, 0, 0, e.shape[0], 1, lay.frame_start[f], 0))
    if pieces.duration_embeds.shape[1]:
        segs.append((pieces.duration_embeds, pieces.duration_embeds.shape[1] * D, 0, pieces.duration_embeds.shape[1], 1,
                     lay.duration_start, 0))
    if pieces.prompt_embeds.shape[1]:
        segs.append((pieces.prompt_embeds, pieces.prompt_embeds.shape[1] * D, 0, pieces.prompt_embeds.shape[1], 1,
                     lay.prompt_start, 0))
    return segs


def _as_bf16(t: torch.Tensor, dev) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.bfloat16).contiguous()


def normalise_pieces(pieces: PromptPieces, dev) -> PromptPieces:
    """bf16, contiguous, on ``dev`` (what the copy kernel reads); masks int64."""
    def m64(t):
        return None if t is None else torch.as_tensor(t).to(device=dev, dtype=torch.long)
    return PromptPieces(
        cue_embeds={m: _as_bf16(t, dev) for m, t in pieces.cue_embeds.items()},
        cue_atts={m: m64(t) for m, t in pieces.cue_atts.items()},
        duration_embeds=_as_bf16(pieces.duration_embeds, dev), duration_atts=m64(pieces.duration_atts),
        prompt_embeds=_as_bf16(pieces.prompt_embeds, dev), prompt_atts=m64(pieces.prompt_atts),
        timestamp_embeds=None if pieces.timestamp_embeds is None else _as_bf16(pieces.timestamp_embeds, dev),
        timestamp_atts=m64(pieces.timestamp_atts),
        enumeration_embeds=None if pieces.enumeration_embeds is None else [_as_bf16(e, dev) for e in pieces.enumeration_embeds])


def copy_pieces(inputs_embeds: torch.Tensor, pieces: PromptPieces, lay: PromptLayout,
                dense_llm: Optional[Dict[str, torch.Tensor]] = None) -> int:
    """Enqueue the copy of every LLM-token piece (and, if given, of dense projected tokens) into ``inputs_embeds``.
    Returns the number of kernel launches."""
    if not inputs_embeds.is_cuda:
        raise _lib.MraError("prompt assembly needs CUDA tensors (no CPU fallback)")
    segs = _segments(pieces, lay, dense_llm)
    launches = 0
    for i in range(0, len(segs), _lib.MAX_PROMPT_SEGMENTS):
        chunk = segs[i:i + _lib.MAX_PROMPT_SEGMENTS]
        arr = (_Seg * len(chunk))()
        for j, (src, vs, fs, rows, frames, dst, dfr) in enumerate(chunk):
            arr[j] = _Seg(src.data_ptr(), vs, fs, rows, frames, dst, dfr)
        check(lib.mra_prompt_assemble(inputs_embeds.data_ptr(), lay.bs, lay.L, lay.D, arr, len(chunk), _lib.current_stream()))
        launches += 1
    return launches


def attention_mask(pieces: PromptPieces, lay: PromptLayout, dev) -> torch.Tensor:
    """``torch.cat(att_list, dim=1)`` of :342-386 -- int64 ``[bs, L]`` (a few KB: plain tensor ops)."""
    bs = lay.bs
    ones_q = torch.ones(bs, lay.num_query, dtype=torch.long, device=dev)
    parts = []
    for f in range(lay.frames):
        if lay.enum_len[f]:
            parts.append(torch.ones(bs, lay.enum_len[f], dtype=torch.long, device=dev))
        for m in lay.modalities:
            parts.append(pieces.cue_atts[m].view(1, -1).expand(bs, -1))
            parts.append(ones_q)
        if lay.ts_len:
            parts.append(pieces.timestamp_atts[:, f, :])
    parts.append(pieces.duration_atts)
    parts.append(pieces.prompt_atts)
    return torch.cat(parts, dim=1)


def targets_with_prefix(text_targets: torch.Tensor, lay: PromptLayout) -> torch.Tensor:
    """:579-588: no loss on the multimodal prefix (``empty_targets`` = -100) followed by the text targets."""
    prefix = torch.full((lay.bs, lay.prompt_start), -100, dtype=torch.long, device=text_targets.device)
    return torch.cat([prefix, text_targets], dim=1)


def query_slot_view(inputs_embeds: torch.Tensor, lay: PromptLayout, modality: str) -> torch.Tensor:
    """``[bs, F, Nq, D]`` strided view of the slots ``llm_proj`` writes for ``modality`` (uniform layouts only)."""
    D = lay.D
    off = (lay.body_start(0) + lay.query_off[modality]) * D
    return inputs_embeds.as_strided((lay.bs, lay.frames, lay.num_query, D), (lay.L * D, lay.frame_stride * D, D, 1),
                                    inputs_embeds.storage_offset() + off)
