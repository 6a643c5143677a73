"""Host logic of the LLM-prompt assembly (no GPU): the row offsets PromptLayout computes are where the reference's
torch.cat (oracle/prompt_oracle.py, models/xinstructblip.py:342-386) puts every piece; the attention mask / targets match."""
import pytest
import torch


def _pieces(bs, F, D, ts, enum, cue=(3, 2), Td=2, Tp=5):
    from mraudio_b200.prompt import PromptPieces
    n = [0]

    def tag(*shape):   # every row gets a distinct value, so positions can be compared
        rows = 1
        for s in shape[:-1]:
            rows *= s
        t = (torch.arange(rows, dtype=torch.float32) + n[0]).view(*shape[:-1], 1).expand(*shape).clone()
        n[0] += rows
        return t
    p = PromptPieces(cue_embeds={"video": tag(cue[0], D), "audio": tag(cue[1], D)},
                     cue_atts={"video": torch.ones(cue[0], dtype=torch.long), "audio": torch.ones(cue[1], dtype=torch.long)},
                     duration_embeds=tag(bs, Td, D), duration_atts=torch.ones(bs, Td, dtype=torch.long),
                     prompt_embeds=tag(bs, Tp, D), prompt_atts=torch.ones(bs, Tp, dtype=torch.long))
    p.prompt_atts[-1, -2:] = 0
    if ts:
        p.timestamp_embeds, p.timestamp_atts = tag(bs, F, ts, D), torch.ones(bs, F, ts, dtype=torch.long)
        p.timestamp_atts[0, F - 1, ts - 1] = 0
    if enum:
        p.enumeration_embeds = [tag(k, D) for k in enum]
    return p, tag


@pytest.mark.parametrize("bs,F,ts,enum,uniform", [(2, 3, 2, None, True), (1, 1, 0, None, True), (2, 4, 1, [4, 3, 3, 3], True),
                                                  (2, 4, 0, [2, 3, 4, 3], False), (3, 2, 2, [1, 5], True)])
def test_layout_offsets_match_reference_concat(bs, F, ts, enum, uniform):
    from mraudio_b200 import prompt as P
    from oracle import prompt_oracle as po
    D, Nq = 8, 32
    pieces, tag = _pieces(bs, F, D, ts, enum)
    inputs_llm = {m: tag(bs, F * Nq, D) for m in ("video", "audio")}
    atts_llm = {m: torch.ones(bs, F * Nq, dtype=torch.long) for m in inputs_llm}
    ref, ref_att, prefix = po.assemble(inputs_llm, atts_llm, pieces.cue_embeds, pieces.cue_atts, pieces.duration_embeds,
                                       pieces.duration_atts, pieces.prompt_embeds, pieces.prompt_atts, pieces.timestamp_embeds,
                                       pieces.timestamp_atts, pieces.enumeration_embeds)
    lay = P.PromptLayout.build(pieces, bs, F, Nq, ("audio", "video"))
    assert lay.L == ref.shape[1] and lay.prompt_start == prefix and lay.uniform == uniform
    # replay the segment table on the CPU (what mra_prompt_assemble does on the GPU) and compare with the concat
    out = torch.full((bs, lay.L, D), -1.0)
    dense = {m: inputs_llm[m].view(bs, F, Nq, D) for m in inputs_llm}
    for src, vs, fs, rows, frames, dst, dfr in P._segments(pieces, lay, dense):
        flat = src.reshape(-1) if src.is_contiguous() else None
        for b in range(bs):
            for f in range(frames):
                if flat is not None:
                    o = b * vs + f * fs
                    blk = flat[o:o + rows * D].view(rows, D)
                else:   # a [:, f] view of a [bs, F, rows, D] piece (non-uniform layouts)
                    blk = src[b]
                out[b, dst + f * dfr: dst + f * dfr + rows] = blk
    assert torch.equal(out, ref)
    assert torch.equal(P.attention_mask(pieces, lay, "cpu"), ref_att)
    if lay.uniform:
        for m in inputs_llm:   # the strided slots llm_proj's epilogue writes
            v = P.query_slot_view(out, lay, m)
            assert torch.equal(v, dense[m])
    tt = torch.arange(bs * pieces.prompt_embeds.shape[1]).view(bs, -1)
    assert torch.equal(P.targets_with_prefix(tt, lay), po.targets(tt, prefix))
