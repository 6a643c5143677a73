"""LLM-prompt assembly: the scatter epilogue of llm_proj + the piece-copy kernel against the CPU
restatement of models/xinstructblip.py:342-386 / :544-594 (oracle/prompt_oracle.py).  Bit-exact: the text pieces are
copies, the query tokens are the same GEMM written through a different TMA tensor map."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(layers=2):
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(3)
    return XInstructBLIPQFormers(modalities=("video", "audio"), num_hidden_layers=layers).cuda().eval()


def _pieces(bs, F, D, g, ts=3, enum=None, cue=(5, 4), Td=2, Tp=17, pad=True):
    from mraudio_b200.prompt import PromptPieces
    r = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
    p = PromptPieces(
        cue_embeds={"video": r(cue[0], D), "audio": r(cue[1], D)},
        cue_atts={"video": torch.ones(cue[0], dtype=torch.long), "audio": torch.ones(cue[1], dtype=torch.long)},
        duration_embeds=r(bs, Td, D), duration_atts=torch.ones(bs, Td, dtype=torch.long),
        prompt_embeds=r(bs, Tp, D), prompt_atts=torch.ones(bs, Tp, dtype=torch.long))
    if pad:   # right-padded prompts / timestamps as the tokenizer produces them
        p.prompt_atts[0, Tp - 3:] = 0
    if ts:
        p.timestamp_embeds, p.timestamp_atts = r(bs, F, ts, D), torch.ones(bs, F, ts, dtype=torch.long)
        if pad:
            p.timestamp_atts[:, 0, ts - 1] = 0
    if enum is not None:
        p.enumeration_embeds = [r(n, D) for n in enum]
    return p


@pytest.mark.parametrize("bs,F,ts,enum", [(2, 3, 3, None), (1, 1, 0, None), (3, 4, 2, [4, 3, 3, 3]), (2, 4, 1, [2, 3, 4, 3]),
                                          (2, 8, 4, None)])
def test_prompt_assembly_matches_reference_concat(bs, F, ts, enum):
    from oracle import prompt_oracle as po
    model = _model()
    D = model.llm_hidden_size
    g = torch.Generator().manual_seed(bs * 10 + F)
    feats = {"video": torch.randn(bs, F, 257, 1408, generator=g).to(torch.bfloat16).cuda(),
             "audio": torch.randn(bs, F, 256, 768, generator=g).to(torch.bfloat16).cuda()}
    ids = torch.randint(1000, 30000, (bs, 8), generator=g).cuda()
    mask = torch.ones(bs, 8, dtype=torch.long).cuda()
    pieces = _pieces(bs, F, D, g, ts=ts, enum=enum)
    with torch.no_grad():
        inputs_llm, atts_llm = model.encode_modalities(feats, ids, mask)
        emb, att = model.encode_modalities(feats, ids, mask, prompt=pieces)
        emb2, att2 = model.encode_modalities(feats, ids, mask, prompt=pieces, scatter_epilogue=False)
    lay = model.last_prompt_layout
    ref_emb, ref_att, prefix = po.assemble({m: v.cpu() for m, v in inputs_llm.items()}, {m: v.cpu() for m, v in atts_llm.items()},
                                           pieces.cue_embeds, pieces.cue_atts, pieces.duration_embeds, pieces.duration_atts,
                                           pieces.prompt_embeds, pieces.prompt_atts, pieces.timestamp_embeds,
                                           pieces.timestamp_atts, pieces.enumeration_embeds)
    assert emb.shape == ref_emb.shape and lay.L == ref_emb.shape[1] and lay.prompt_start == prefix
    assert torch.equal(emb.cpu(), ref_emb), "scatter-epilogue assembly differs from the reference concat"
    assert torch.equal(emb2.cpu(), ref_emb), "copy assembly differs from the reference concat"
    assert torch.equal(att.cpu(), ref_att) and torch.equal(att2.cpu(), ref_att)
    from mraudio_b200.prompt import targets_with_prefix
    tt = torch.randint(0, 32000, (bs, pieces.prompt_embeds.shape[1]), generator=g)
    assert torch.equal(targets_with_prefix(tt.cuda(), lay).cpu(), po.targets(tt, prefix))


def test_scatter_epilogue_is_used_and_saves_the_copy():
    model = _model()
    g = torch.Generator().manual_seed(5)
    bs, F = 2, 4
    feats = {"video": torch.randn(bs, F, 257, 1408, generator=g).to(torch.bfloat16).cuda(),
             "audio": torch.randn(bs, F, 256, 768, generator=g).to(torch.bfloat16).cuda()}
    ids = torch.randint(1000, 30000, (bs, 8), generator=g).cuda()
    mask = torch.ones(bs, 8, dtype=torch.long).cuda()
    pieces = _pieces(bs, F, model.llm_hidden_size, g)
    with torch.no_grad():
        model.encode_modalities(feats, ids, mask)
        base = model.last_launches
        model.encode_modalities(feats, ids, mask, prompt=pieces)
        assert model.last_launches == base + 1      # one piece-copy launch; the query tokens cost no extra launch


def test_prompt_segment_bounds_are_checked():
    from mraudio_b200 import _lib
    from mraudio_b200.prompt import _Seg
    out = torch.zeros(2, 10, 64, dtype=torch.bfloat16, device="cuda")
    src = torch.zeros(2, 4, 64, dtype=torch.bfloat16, device="cuda")
    arr = (_Seg * 1)(_Seg(src.data_ptr(), 4 * 64, 0, 4, 1, 8, 0))     # rows 8..11 of a 10-row sequence
    assert _lib.lib.mra_prompt_assemble(out.data_ptr(), 2, 10, 64, arr, 1, _lib.current_stream()) != 0
    assert b"exceeds" in _lib.lib.mra_last_error()
