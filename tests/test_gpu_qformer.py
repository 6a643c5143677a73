"""GPU parity of the whole Q-Former + llm_proj forward (through mraudio_b200.qformer -> C-ABI) against the CPU oracle
and the golden fixtures frozen from the HF port of the LAVIS Q-Former.

Tolerance (BASELINE.json north_star): max|y - y_ref| / max|y_ref| <= 2e-2 against the fp32 reference.  Against the
bf16-storage emulation of the oracle (same rounding points as the CUDA path) the bound is much tighter."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import qformer_oracle as qo

pytestmark = pytest.mark.gpu
TOL_FP32_REF = 2e-2
TOL_EMULATED = 6e-3


def _rel(a, b):
    return ((a.float().cpu() - b.float()).abs().max() / b.float().abs().max()).item()


def _build(cfg: qo.QFormerOracleConfig, w, llm_dim=4096):
    from mraudio_b200 import BertConfig, BertLMHeadModel, LLMProjB200
    bc = BertConfig.from_pretrained("bert-base-uncased")
    bc.encoder_width = cfg.encoder_width
    bc.add_cross_attention = True
    bc.cross_attention_freq = cfg.cross_attention_freq
    bc.query_length = cfg.query_length
    bc.num_hidden_layers = cfg.num_hidden_layers
    bc.vocab_size = cfg.vocab_size
    model = BertLMHeadModel(bc)
    sd = {k: v for k, v in w.items() if k.startswith("bert.")}
    msg = model.load_state_dict(sd, strict=False)
    assert not msg.missing_keys and not msg.unexpected_keys, msg
    proj = LLMProjB200(cfg.hidden_size, llm_dim)
    proj.load_state_dict({"weight": w["llm_proj.weight"], "bias": w["llm_proj.bias"]})
    return model.cuda().eval(), proj.cuda().eval()


@pytest.mark.parametrize("name,W,Nk", [("video", 1408, 257), ("audio", 768, 256)])
def test_forward_matches_golden_and_oracle(name, W, Nk):
    fx = np.load(os.path.join(GOLDEN, f"qformer_{name}.npz"))
    cfg = qo.QFormerOracleConfig(encoder_width=W)
    w = qo.init_qformer_weights(cfg, seed=int(fx["weight_seed"]), randomize_ln_and_bias=True)
    g = torch.Generator().manual_seed(int(fx["input_seed"]))
    ids = torch.randint(1000, 30000, (2, 32), generator=g)
    enc = torch.randn(2, Nk, W, generator=g)
    tmask = torch.from_numpy(fx["text_mask"])
    atts = torch.cat([torch.ones(2, 32, dtype=torch.long), tmask], 1)
    model, proj = _build(cfg, w)
    qe = w["query_tokens"].expand(2, -1, -1).cuda()
    with torch.no_grad():
        out = model.bert(ids.cuda(), attention_mask=atts.cuda(), query_embeds=qe, encoder_hidden_states=enc.cuda(),
                         encoder_attention_mask=torch.ones(2, Nk, dtype=torch.long).cuda(), return_dict=True)
        hid = out.last_hidden_state
        y = proj(hid[:, :32, :])
    torch.cuda.synchronize()
    assert hid.shape == (2, 64, 768) and y.shape == (2, 32, 4096)
    ref = torch.from_numpy(fx["last_hidden_state"])
    # padded text positions of row 1 are don't-care in the reference too (they only attend, nobody attends to them)
    assert _rel(hid[:, :32], ref[:, :32]) < TOL_FP32_REF
    assert _rel(hid[0], ref[0]) < TOL_FP32_REF
    assert _rel(y[:, :, ::64], torch.from_numpy(fx["llm_proj_sample"])) < TOL_FP32_REF
    with torch.no_grad():
        emu = qo.qformer_bert(w, cfg, ids, atts, w["query_tokens"], enc.to(torch.bfloat16).float(), None, emulate_bf16=True)
    assert _rel(hid[:, :32], emu[:, :32]) < TOL_EMULATED


def test_query_only_two_layer_qformer_1024_keys():
    """Video-LLaMA-v1-style video Q-Former shape (parity unpinned w.r.t. the reference; pinned to HF Blip2QFormerModel)."""
    fx = np.load(os.path.join(GOLDEN, "qformer_queryonly.npz"))
    cfg = qo.QFormerOracleConfig(encoder_width=768, num_hidden_layers=2, cross_attention_freq=1, has_text=False)
    w = qo.init_qformer_weights(cfg, seed=int(fx["weight_seed"]), randomize_ln_and_bias=True)
    g = torch.Generator().manual_seed(int(fx["input_seed"]))
    enc = torch.randn(2, 1024, 768, generator=g)
    model, _ = _build(cfg, w)
    with torch.no_grad():
        hid = model.bert(query_embeds=w["query_tokens"].cuda(), encoder_hidden_states=enc.cuda(), return_dict=True).last_hidden_state
    assert _rel(hid, torch.from_numpy(fx["last_hidden_state"])) < TOL_FP32_REF


@pytest.mark.parametrize("rows,T,Nk,W,layers", [(5, 32, 257, 1408, 4), (3, 0, 64, 768, 2), (7, 11, 19, 768, 3),
                                                 (130, 32, 257, 1408, 2)])
def test_forward_shapes_fused_projection_and_dead_ffn_skip(rows, T, Nk, W, layers):
    cfg = qo.QFormerOracleConfig(encoder_width=W, num_hidden_layers=layers, has_text=T > 0)
    w = qo.init_qformer_weights(cfg, seed=rows, randomize_ln_and_bias=True, llm_dim=512)
    g = torch.Generator().manual_seed(rows + 1)
    enc = torch.randn(rows, Nk, W, generator=g).to(torch.bfloat16)
    ids = torch.randint(1000, 30000, (rows, T), generator=g) if T else None
    atts = None
    if T:
        tmask = torch.ones(rows, T, dtype=torch.long)
        tmask[0, T // 2:] = 0
        atts = torch.cat([torch.ones(rows, 32, dtype=torch.long), tmask], 1)
    model, proj = _build(cfg, w, llm_dim=512)
    with torch.no_grad():
        out = model.bert(ids.cuda() if T else None, attention_mask=atts.cuda() if T else None,
                         query_embeds=w["query_tokens"].cuda(), encoder_hidden_states=enc.cuda(), return_dict=True,
                         llm_proj=proj, skip_dead_text_ffn=True)
        ref = qo.qformer_bert(w, cfg, ids, atts, w["query_tokens"], enc.float(), None, emulate_bf16=True,
                              skip_dead_text_ffn=True)
        ref32 = qo.qformer_bert(w, cfg, ids, atts, w["query_tokens"], enc.float(), None)
        ref_y = qo.llm_proj(w, ref32[:, :32])
    assert _rel(out.last_hidden_state[:, :32], ref[:, :32]) < TOL_EMULATED
    assert _rel(out.last_hidden_state[:, :32], ref32[:, :32]) < TOL_FP32_REF
    assert out.llm_inputs.shape == (rows, 32, 512)
    assert _rel(out.llm_inputs, ref_y) < TOL_FP32_REF
    # two-call form == fused form
    with torch.no_grad():
        y2 = proj(out.last_hidden_state[:, :32, :])
    assert _rel(y2, ref_y) < TOL_FP32_REF


def test_split_residual_stream_matches_fp32_residual_stream(monkeypatch):
    """The inference forward keeps the residual stream as a bf16 (hi, lo) pair (16 mantissa bits, csrc/gemm_ln.cu SPLIT);
    MRA_SPLIT_RESIDUAL=0 selects the fp32 stream.  Same inputs: both inside the tolerance against the fp32 oracle and
    within the bf16-emulation bound of each other over 12 layers (the split form loses only the bits below 2^-17 of every
    residual, but any perturbation flips some bf16 roundings of the GEMM operands, which then propagate: measured 2.2e-3)."""
    cfg = qo.QFormerOracleConfig(encoder_width=1408, num_hidden_layers=12)
    w = qo.init_qformer_weights(cfg, seed=5, randomize_ln_and_bias=True, llm_dim=512)
    g = torch.Generator().manual_seed(6)
    rows, T, Nk = 9, 32, 257
    enc = torch.randn(rows, Nk, 1408, generator=g).to(torch.bfloat16)
    ids = torch.randint(1000, 30000, (rows, T), generator=g)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MRA_SPLIT_RESIDUAL", mode)
        model, proj = _build(cfg, w, llm_dim=512)
        with torch.no_grad():
            o = model.bert(ids.cuda(), query_embeds=w["query_tokens"].cuda(), encoder_hidden_states=enc.cuda(), return_dict=True,
                           llm_proj=proj)
        outs[mode] = (o.last_hidden_state.float().cpu(), o.llm_inputs.float().cpu())
    with torch.no_grad():
        ref32 = qo.qformer_bert(w, cfg, ids, None, w["query_tokens"], enc.float(), None)
        ref_y = qo.llm_proj(w, ref32[:, :32])
    for mode in ("1", "0"):
        assert _rel(outs[mode][0], ref32) < TOL_FP32_REF, mode
        assert _rel(outs[mode][1], ref_y) < TOL_FP32_REF, mode
    assert _rel(outs["1"][0], outs["0"][0]) < TOL_EMULATED
    assert not torch.equal(outs["1"][0], outs["0"][0])   # the two modes really ran different kernels


@pytest.mark.parametrize("rows", [3, 10])
def test_fused_qkv_attention_forward_equals_unfused_forward(monkeypatch, rows):
    """The inference forward runs the QKV Linear + self-attention core as one kernel when every row has 32 + 32 tokens
    (csrc/qkv_attn.cu) and hands head-major Q / K / V to the attention kernel otherwise (MRA_FUSE_QKV_ATTN=0; row-major with
    MRA_HEAD_MAJOR=0).  All three forms compute identical values: the outputs are bit-equal."""
    cfg = qo.QFormerOracleConfig(encoder_width=1408, num_hidden_layers=4)
    w = qo.init_qformer_weights(cfg, seed=7, randomize_ln_and_bias=True, llm_dim=512)
    g = torch.Generator().manual_seed(8)
    T, Nk = 32, 257
    enc = torch.randn(rows, Nk, 1408, generator=g).to(torch.bfloat16)
    ids = torch.randint(1000, 30000, (rows, T), generator=g)
    atts = torch.ones(rows, 32 + T, dtype=torch.long)
    atts[1, 32 + 20:] = 0
    outs, launches = {}, {}
    for mode in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("MRA_FUSE_QKV_ATTN", mode[0])
        monkeypatch.setenv("MRA_HEAD_MAJOR", mode[1])
        model, proj = _build(cfg, w, llm_dim=512)
        with torch.no_grad():
            o = model.bert(ids.cuda(), attention_mask=atts.cuda(), query_embeds=w["query_tokens"].cuda(),
                           encoder_hidden_states=enc.cuda(), return_dict=True, llm_proj=proj)
        outs[mode] = (o.last_hidden_state.float().cpu(), o.llm_inputs.float().cpu())
        launches[mode] = model.bert.last_launches
    assert launches[("1", "1")] == launches[("0", "1")] - cfg.num_hidden_layers    # one launch less per layer
    for mode in (("0", "1"), ("0", "0")):
        assert torch.equal(outs[mode][0], outs[("1", "1")][0]) and torch.equal(outs[mode][1], outs[("1", "1")][1]), mode
    with torch.no_grad():
        ref32 = qo.qformer_bert(w, cfg, ids, atts, w["query_tokens"], enc.float(), None)
    assert _rel(outs[("1", "1")][0][:, :32], ref32[:, :32]) < TOL_FP32_REF


def test_host_pipeline_streams_batches_through_reused_slots():
    """``HostPipeline`` (pinned host features -> H2D -> both Q-Formers -> D2H into pinned host buffers, three streams, two
    slots whose device buffers are reused without per-batch allocation): five different batches submitted back to back --
    more than there are slots, the host far ahead of the device -- each come out equal to ``encode_modalities`` on the same
    batch, and nothing is allocated on the device after the first round through the slots."""
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    widths = {"video": 1408, "audio": 768}
    model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features=widths, llm_hidden_size=512,
                                  num_hidden_layers=2).cuda().eval()
    bs, Fr, T = 2, 4, 32
    tokens = {"video": 257, "audio": 64}
    pipe = model.host_pipeline(bs, Fr, tokens, T, slots=2)
    g = torch.Generator().manual_seed(9)
    batches = []
    for i in range(5):
        feats = {m: torch.randn(bs, Fr, tokens[m], widths[m], generator=g).to(torch.bfloat16).pin_memory() for m in tokens}
        ids = torch.randint(1000, 30000, (bs, T), generator=g).pin_memory()
        mask = torch.ones(bs, T, dtype=torch.long)
        mask[i % bs, 7 + i:] = 0
        batches.append((feats, ids, mask.pin_memory()))
    outs, allocated = [], []
    for feats, ids, mask in batches:
        sl = pipe.submit(feats, ids, mask)
        allocated.append(torch.cuda.memory_allocated())
        # a consumer reads a slot's pinned output once that slot's D2H event has completed (before the slot comes round again)
        sl.out_done.synchronize()
        outs.append({m: t.clone() for m, t in sl.out_host.items()})
    assert allocated[2] == allocated[3] == allocated[4]          # steady state: no device allocation per batch
    assert pipe.h2d_bytes == sum(t.numel() * 2 for t in batches[0][0].values()) + 2 * bs * T * 8
    assert pipe.d2h_bytes == 2 * bs * Fr * 32 * 512 * 2
    for (feats, ids, mask), out in zip(batches, outs):
        with torch.no_grad():
            ref, _ = model.encode_modalities({m: t.cuda() for m, t in feats.items()}, ids.cuda(), mask.cuda())
        for m in ref:
            assert torch.equal(out[m], ref[m].cpu()), m
    # without waiting in between: the slots' events alone order reuse (batch i + 2 overwrites slot i only after its D2H)
    last = [pipe.submit(*b) for b in batches]
    torch.cuda.synchronize()
    for m in outs[4]:
        assert torch.equal(last[4].out_host[m], outs[4][m]) and torch.equal(last[3].out_host[m], outs[3][m])


def test_forward_is_deterministic_and_linear_in_projection():
    cfg = qo.QFormerOracleConfig(encoder_width=768, num_hidden_layers=2)
    w = qo.init_qformer_weights(cfg, seed=0, llm_dim=256)
    model, proj = _build(cfg, w, llm_dim=256)
    g = torch.Generator().manual_seed(0)
    enc = torch.randn(4, 50, 768, generator=g).cuda()
    ids = torch.randint(1000, 30000, (4, 8), generator=g).cuda()
    kw = dict(query_embeds=w["query_tokens"].cuda(), encoder_hidden_states=enc, return_dict=True, llm_proj=proj)
    with torch.no_grad():
        a = model.bert(ids, **kw)
        b = model.bert(ids, **kw)
    assert torch.equal(a.last_hidden_state, b.last_hidden_state) and torch.equal(a.llm_inputs, b.llm_inputs)
    # rows are independent: permuting the batch permutes the output (frames folded into the batch dimension)
    perm = torch.tensor([2, 0, 3, 1]).cuda()
    with torch.no_grad():
        c = model.bert(ids[perm], **dict(kw, encoder_hidden_states=enc[perm]))
    assert torch.equal(c.last_hidden_state, a.last_hidden_state[perm])


def test_errors_are_loud():
    from mraudio_b200 import MraError
    cfg = qo.QFormerOracleConfig(encoder_width=768, num_hidden_layers=1)
    w = qo.init_qformer_weights(cfg, seed=0, llm_dim=64)
    model, _ = _build(cfg, w, llm_dim=64)
    with pytest.raises(MraError):
        model.bert(query_embeds=w["query_tokens"], encoder_hidden_states=torch.zeros(1, 4, 768))  # CPU tensors
    with pytest.raises(ValueError):
        model.bert(query_embeds=w["query_tokens"].cuda(), encoder_hidden_states=torch.zeros(1, 4, 1408).cuda())


def test_xinstructblip_encode_modalities_lockstep_matches_oracle_and_separate_calls():
    """models/xinstructblip.py:280-305 through the host mirror: both modalities in one lockstep call (grouped GEMM
    launches) == one call per modality == oracle, including the reference's frame-major text tiling for bs > 1."""
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    widths = {"video": 1408, "audio": 768}
    model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features=widths, llm_hidden_size=512,
                                  num_hidden_layers=3).cuda().eval()
    g = torch.Generator().manual_seed(4)
    bs, Fr, T = 3, 2, 9
    feats = {"video": torch.randn(bs, Fr, 257, 1408, generator=g).to(torch.bfloat16),
             "audio": torch.randn(bs, Fr, 40, 768, generator=g).to(torch.bfloat16)}
    ids = torch.randint(1000, 30000, (bs, T), generator=g)
    mask = torch.ones(bs, T, dtype=torch.long)
    mask[1, 5:] = 0
    dfe = {m: t.cuda() for m, t in feats.items()}
    with torch.no_grad():
        a, atts = model.encode_modalities(dfe, ids.cuda(), mask.cuda())
        n_lock = model.last_launches
        model.lockstep_modalities = False
        b, _ = model.encode_modalities(dfe, ids.cuda(), mask.cuda())
        n_sep = model.last_launches
        # the reference's per-frame list form gives the same rows (frame fold + batch-major reorder, :280-285)
        lists = {m: [t[:, f].contiguous() for f in range(Fr)] for m, t in dfe.items()}
        c, _ = model.encode_modalities(lists, ids.cuda(), mask.cuda())
    assert n_lock < n_sep
    for m in ("video", "audio"):
        assert a[m].shape == (bs, Fr * 32, 512) and atts[m].shape == (bs, Fr * 32)
        assert torch.equal(a[m], b[m]) and torch.equal(a[m], c[m])
        cfg = qo.QFormerOracleConfig(encoder_width=widths[m], num_hidden_layers=3)
        w = {"bert." + k[len(f"{m}_Qformer.bert."):]: v.detach().cpu().float() for k, v in model.state_dict().items()
             if k.startswith(f"{m}_Qformer.bert.")}
        w["query_tokens"] = getattr(model, f"{m}_query_tokens").detach().cpu()
        w["llm_proj.weight"] = getattr(model, f"{m}_llm_proj").weight.detach().cpu()
        w["llm_proj.bias"] = getattr(model, f"{m}_llm_proj").bias.detach().cpu()
        with torch.no_grad():
            ref = qo.xinstructblip_encode(w, cfg, feats[m].float(), ids, mask)
            other = qo.xinstructblip_encode(w, cfg, feats[m].float(), ids, mask, match_reference_text_tiling=False)
        assert _rel(a[m], ref) < TOL_FP32_REF
        assert _rel(a[m], other) > _rel(a[m], ref)      # the "fixed" tiling is NOT what the reference computes


def _oracle_weights(model, m):
    w = {"bert." + k[len(f"{m}_Qformer.bert."):]: v.detach().cpu().float() for k, v in model.state_dict().items()
         if k.startswith(f"{m}_Qformer.bert.")}
    w["query_tokens"] = getattr(model, f"{m}_query_tokens").detach().cpu()
    w["llm_proj.weight"] = getattr(model, f"{m}_llm_proj").weight.detach().cpu()
    w["llm_proj.bias"] = getattr(model, f"{m}_llm_proj").bias.detach().cpu()
    return w


@pytest.mark.parametrize("bs,Fr,T", [(32, 8, 32), (4, 75, 32)], ids=["config2_32x8", "config5_shape_4x75"])
def test_full_size_configs_row_independence_and_sampled_oracle_parity(bs, Fr, T):
    """BASELINE.json's full sizes (config 2: 32 videos x 8 frames; config 5's 75-clip videos), 12 layers, both modalities in
    lockstep.  The oracle cannot run 256 rows x 12 layers in seconds, so parity is shown through (a) a size-independent
    property -- every (video, frame) row is independent of the rest of the batch: the rows of two videos computed inside
    the full batch equal the same two videos computed alone -- and (b) the fp32 oracle on exactly those two videos."""
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio")).cuda().eval()
    g = torch.Generator().manual_seed(bs + Fr)
    feats = {"video": torch.randn(bs, Fr, 257, 1408, generator=g).to(torch.bfloat16),
             "audio": torch.randn(bs, Fr, 256, 768, generator=g).to(torch.bfloat16)}
    ids = torch.randint(1000, 30000, (bs, T), generator=g)
    mask = torch.ones(bs, T, dtype=torch.long)
    mask[1, T - 5:] = 0
    pick = [1, bs - 1]
    # frame-major text tiling (:287-289) pairs row b*F+f with text (b*F+f) % bs: a sub-batch sees other texts, so the
    # independence check uses the batch-major pairing (each video with its own prompt)
    with torch.no_grad():
        full, _ = model.encode_modalities({m: t.cuda() for m, t in feats.items()}, ids.cuda(), mask.cuda(),
                                          match_reference_text_tiling=False)
        sub, _ = model.encode_modalities({m: t[pick].cuda() for m, t in feats.items()}, ids[pick].cuda(), mask[pick].cuda(),
                                         match_reference_text_tiling=False)
    for m in ("video", "audio"):
        assert full[m].shape == (bs, Fr * 32, 4096) and torch.isfinite(full[m].float()).all()
        assert torch.equal(full[m][pick], sub[m]), f"{m}: rows depend on the rest of the batch"
    nf = min(Fr, 4)     # oracle on the first frames of the two videos (fp32 CPU, 12 layers)
    for m in ("video", "audio"):
        cfg = qo.QFormerOracleConfig(encoder_width=feats[m].shape[-1])
        with torch.no_grad():
            ref = qo.xinstructblip_encode(_oracle_weights(model, m), cfg, feats[m][pick][:, :nf].float(), ids[pick], mask[pick],
                                          match_reference_text_tiling=False)
        got = full[m][pick].view(2, Fr, 32, 4096)[:, :nf].reshape(2, nf * 32, 4096)
        assert _rel(got, ref) < TOL_FP32_REF, (m, _rel(got, ref))


def test_maximum_text_length_and_single_row():
    """Edge sizes: T = 128 prompt tokens (the reference's max_txt_len) with ragged padding, and a single (video, frame) row."""
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(1)
    model = XInstructBLIPQFormers(modalities=("video",), num_hidden_layers=2, llm_hidden_size=256).cuda().eval()
    g = torch.Generator().manual_seed(9)
    for bs, Fr, T in ((2, 2, 128), (1, 1, 5)):
        feats = {"video": torch.randn(bs, Fr, 257, 1408, generator=g).to(torch.bfloat16)}
        ids = torch.randint(1000, 30000, (bs, T), generator=g)
        mask = torch.ones(bs, T, dtype=torch.long)
        mask[0, T // 3:] = 0
        with torch.no_grad():
            out, _ = model.encode_modalities({"video": feats["video"].cuda()}, ids.cuda(), mask.cuda())
            cfg = qo.QFormerOracleConfig(encoder_width=1408, num_hidden_layers=2)
            ref = qo.xinstructblip_encode(_oracle_weights(model, "video"), cfg, feats["video"].float(), ids, mask)
        assert _rel(out["video"], ref) < TOL_FP32_REF
