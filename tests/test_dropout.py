"""Training-mode dropout (reference: model.train() at utils/trainer.py:110 with the Q-Former's BertConfig
hidden_dropout_prob = attention_probs_dropout_prob = 0.1; HF port modeling_instructblip.py:530,551,608,781).

The CUDA path never stores a mask: forward and backward regenerate it from a counter-based Philox4x32-10 stream
(csrc/dropout.cuh) and oracle/qformer_oracle.py holds the same generator, so output and gradients are compared under the
IDENTICAL mask.  CPU part: the generator against the published Philox known-answer vectors and the statistics of the masks."""
import numpy as np
import pytest
import torch

from oracle import qformer_oracle as qo


def test_philox_known_answers_and_mask_statistics():
    # Random123 known-answer vectors of philox4x32-10
    z = np.array([0])
    assert [int(x[0]) for x in qo._philox4x32_10(z, z, z, z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.array([0xFFFFFFFF])
    assert [int(x[0]) for x in qo._philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    pi = qo._philox4x32_10(np.array([0x243f6a88]), np.array([0x85a308d3]), np.array([0x13198a2e]), np.array([0x03707344]), 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in pi] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # mean preservation: E[multiplier] = 1, drop fraction = round(256 p) / 256
    for p in (0.1, 0.25):
        thr, scale = qo.dropout_params(p)
        m = qo.hidden_dropout_mult(p, 99, qo.DROP_FFN_OUT, 3, torch.arange(512).reshape(8, 64), 768)
        n = m.numel()
        drop = (m == 0).double().mean().item()
        sigma = (thr / 256 * (1 - thr / 256) / n) ** 0.5
        assert abs(drop - thr / 256) < 4 * sigma
        assert abs(m.double().mean().item() - 1.0) < 4 * sigma * scale
        assert set(torch.unique(m).tolist()) == {0.0, float(np.float32(scale))}
        a = qo.attention_dropout_mult(p, 99, qo.DROP_CROSS_PROBS, 0, 3, 12, 32, 257)
        assert abs((a == 0).double().mean().item() - thr / 256) < 4 * (thr / 256 * (1 - thr / 256) / a.numel()) ** 0.5
    # different sites / layers / seeds draw different masks; the same arguments the same mask
    m0 = qo.hidden_dropout_mult(0.1, 1, 4, 0, torch.arange(64), 768)
    assert torch.equal(m0, qo.hidden_dropout_mult(0.1, 1, 4, 0, torch.arange(64), 768))
    for other in (qo.hidden_dropout_mult(0.1, 2, 4, 0, torch.arange(64), 768), qo.hidden_dropout_mult(0.1, 1, 5, 0, torch.arange(64), 768),
                  qo.hidden_dropout_mult(0.1, 1, 4, 1, torch.arange(64), 768)):
        assert not torch.equal(m0, other)


def _build(cfg, w, llm_dim):
    from mraudio_b200 import BertConfig, BertLMHeadModel, LLMProjB200
    bc = BertConfig.from_pretrained("bert-base-uncased")
    bc.encoder_width, bc.cross_attention_freq, bc.query_length = cfg.encoder_width, cfg.cross_attention_freq, cfg.query_length
    bc.num_hidden_layers, bc.vocab_size = cfg.num_hidden_layers, cfg.vocab_size
    q = BertLMHeadModel(bc)
    q.load_state_dict({k: v for k, v in w.items() if k.startswith("bert.")}, strict=False)
    proj = LLMProjB200(cfg.hidden_size, llm_dim)
    proj.load_state_dict({"weight": w["llm_proj.weight"], "bias": w["llm_proj.bias"]})
    return q.cuda(), torch.nn.Parameter(w["query_tokens"].clone().cuda()), proj.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("rows,T,Nk,W,layers,p", [(3, 32, 64, 64, 2, 0.1), (2, 32, 257, 1408, 3, 0.1), (4, 0, 40, 768, 2, 0.25),
                                                        (2, 32, 600, 64, 2, 0.1)])
def test_dropout_forward_and_gradients_match_oracle_under_the_same_mask(rows, T, Nk, W, layers, p):
    from mraudio_b200.training import TrainableQFormer
    D, seed = 256, 20261018 + rows
    cfg = qo.QFormerOracleConfig(encoder_width=W, num_hidden_layers=layers, has_text=T > 0)
    w = qo.init_qformer_weights(cfg, seed=rows + layers, llm_dim=D, randomize_ln_and_bias=True)
    g = torch.Generator().manual_seed(7)
    enc = torch.randn(rows, Nk, W, generator=g).to(torch.bfloat16).float()
    ids = torch.randint(1000, 30000, (rows, T), generator=g) if T else None
    atts = None
    if T:
        tm = torch.ones(rows, T, dtype=torch.long)
        tm[0, T // 2:] = 0
        atts = torch.cat([torch.ones(rows, 32, dtype=torch.long), tm], 1)
    G = torch.randn(rows, 32, D, generator=g)
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    hid = qo.qformer_bert(wr, cfg, ids, atts, wr["query_tokens"], enc, None, skip_dead_text_ffn=True, dropout=(p, seed))
    y_ref = qo.llm_proj(wr, hid[:, :32])
    (y_ref * G).sum().backward()
    with torch.no_grad():
        y_eval = qo.llm_proj(w, qo.qformer_bert(w, cfg, ids, atts, w["query_tokens"], enc, None, skip_dead_text_ffn=True)[:, :32])

    q, qt, proj = _build(cfg, w, D)
    st = TrainableQFormer(q, qt, proj)
    args = (enc.cuda(), ids.cuda() if T else None, atts.cuda() if T else None)
    y = st.forward(*args, dropout_p=p, dropout_seed=seed)
    rel = lambda a, b: ((a.float().cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(y, y_ref.detach()) < 2e-2
    assert rel(y, y_eval) > 5e-2                     # the mask really acted: far from the eval-mode output ...
    (y.float() * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    sd = dict(q.named_parameters())
    checked, worst = 0, 0.0
    for k, v in wr.items():
        gr = v.grad if v.grad is not None else torch.zeros_like(v)
        got = sd[k].grad if k.startswith("bert.") else {"query_tokens": qt.grad, "llm_proj.weight": proj.weight.grad,
                                                        "llm_proj.bias": proj.bias.grad}.get(k)
        if got is None or gr.abs().max().item() == 0.0 or k.endswith("attention.self.key.bias"):
            continue
        r = rel(got, gr)
        worst = max(worst, r)
        assert r < 4e-2, (k, r)
        checked += 1
    assert checked > 20
    # ... the same seed reproduces the output bit for bit, another seed does not, p = 0 is the eval-mode forward
    st._saved = None
    y2 = st.forward(*args, dropout_p=p, dropout_seed=seed)
    y3 = st.forward(*args, dropout_p=p, dropout_seed=seed + 1)
    y4 = st.forward(*args)
    assert torch.equal(y, y2) and not torch.equal(y, y3)
    assert rel(y4, y_eval) < 2e-2
    print("worst relative gradient error under dropout", worst)


@pytest.mark.gpu
def test_trainer_with_reference_dropout_pads_text_and_learns():
    """QFormerTrainer(dropout=None) = the reference's model.train(): p from the Q-Former config (0.1); a prompt of 8 tokens is
    padded to 32 masked tokens for the TMA attention kernels; the surrogate loss still goes down."""
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                  llm_hidden_size=128, num_hidden_layers=2).cuda()
    tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-3, dropout=None, seed=3)
    assert model.video_Qformer.config.hidden_dropout_prob == 0.1
    g = torch.Generator().manual_seed(3)
    feats = {"video": torch.randn(2, 3, 17, 128, generator=g).cuda().to(torch.bfloat16),
             "audio": torch.randn(2, 3, 16, 64, generator=g).cuda().to(torch.bfloat16)}
    ids = torch.randint(1000, 30000, (2, 8), generator=g).cuda()
    mask = torch.ones(2, 8, dtype=torch.long).cuda()
    sur = {m: torch.randn(2, 3 * 32, 128, generator=g).cuda() for m in feats}
    losses = [tr.train_step(feats, ids, mask, surrogate=sur).item() for _ in range(10)]
    assert losses[-1] < losses[0], losses
    assert len(set(losses)) == len(losses)
