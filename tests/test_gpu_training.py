"""GPU parity of the Q-Former / projection backward + Adam step (mra_qformer_backward, mra_adam_step through
mraudio_b200.training) against autograd of the fp32 oracle with the surrogate loss L = sum(inputs_llm * G)
(the LLM that produces the real loss is out of scope)."""
import pytest
import torch

from oracle import qformer_oracle as qo

pytestmark = pytest.mark.gpu


def _build(cfg, w, llm_dim):
    from mraudio_b200 import BertConfig, BertLMHeadModel, LLMProjB200
    bc = BertConfig.from_pretrained("bert-base-uncased")
    bc.encoder_width, bc.cross_attention_freq, bc.query_length = cfg.encoder_width, cfg.cross_attention_freq, cfg.query_length
    bc.num_hidden_layers, bc.vocab_size = cfg.num_hidden_layers, cfg.vocab_size
    q = BertLMHeadModel(bc)
    msg = q.load_state_dict({k: v for k, v in w.items() if k.startswith("bert.")}, strict=False)
    assert not msg.missing_keys
    proj = LLMProjB200(cfg.hidden_size, llm_dim)
    proj.load_state_dict({"weight": w["llm_proj.weight"], "bias": w["llm_proj.bias"]})
    qt = torch.nn.Parameter(w["query_tokens"].clone())
    return q.cuda(), torch.nn.Parameter(qt.data.cuda()), proj.cuda()


def _oracle_grads(cfg, w, ids, atts, enc, G):
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    hid = qo.qformer_bert(wr, cfg, ids, atts, wr["query_tokens"], enc, None, skip_dead_text_ffn=True)
    y = qo.llm_proj(wr, hid[:, :cfg.query_length])
    loss = (y * G).sum()
    loss.backward()
    return y.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in wr.items()}


# the last case is BASELINE.json config 4 at its per-GPU size: 8 videos x 8 frames = 64 rows of 257 x 1408 tokens, 12 layers
# (2, 32, 600, 64, 2): four resident key chunks in the attention backward (257 keys: two); (2, 96, 200, 64, 2): 128-token
# self-attention rows (eight 16-query tiles) with 200 keys in two chunks
@pytest.mark.parametrize("rows,T,Nk,W,layers", [(3, 8, 20, 64, 2), (2, 32, 257, 1408, 3), (5, 0, 40, 768, 2),
                                                 (2, 32, 600, 64, 2), (2, 96, 200, 64, 2), (64, 32, 257, 1408, 12)])
def test_backward_matches_oracle_autograd(rows, T, Nk, W, layers):
    from mraudio_b200.training import TrainableQFormer
    D = 256
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
    y_ref, gref = _oracle_grads(cfg, w, ids, atts, enc, G)

    q, qt, proj = _build(cfg, w, D)
    st = TrainableQFormer(q, qt, proj)
    y = st.forward(enc.cuda(), ids.cuda() if T else None, atts.cuda() if T else None)
    assert ((y.float().cpu() - y_ref).abs().max() / y_ref.abs().max()).item() < 2e-2
    (y.float() * G.cuda()).sum().backward()
    torch.cuda.synchronize()

    sd = dict(q.named_parameters())
    checked = 0
    worst = 0.0
    for k, gr in gref.items():
        if k.startswith("bert."):
            got = sd[k].grad
        elif k == "query_tokens":
            got = qt.grad
        elif k == "llm_proj.weight":
            got = proj.weight.grad
        elif k == "llm_proj.bias":
            got = proj.bias.grad
        else:
            continue
        if gr.abs().max().item() == 0.0:
            # unused parameters (dead last-layer text FFN, unused embedding rows, text weights when T == 0): zero grads
            assert got is None or got.abs().max().item() == 0.0, k
            continue
        if k.endswith("attention.self.key.bias"):
            # softmax is invariant to a per-query constant, so dL/d(key bias) is exactly 0 in exact arithmetic: the
            # oracle holds fp32 noise there; require the CUDA value to be noise relative to the query-bias gradient
            ref_scale = gref[k.replace(".key.", ".query.")].abs().max().item()
            assert got.abs().max().item() < 5e-2 * ref_scale, (k, got.abs().max().item(), ref_scale)
            continue
        rel = ((got.float().cpu() - gr).abs().max() / gr.abs().max()).item()
        worst = max(worst, rel)
        assert rel < 4e-2, (k, rel)
        checked += 1
    assert checked > 20
    print("worst relative gradient error", worst)


def test_gradient_accumulation_adam_and_linearity():
    from mraudio_b200.training import TrainableQFormer
    cfg = qo.QFormerOracleConfig(encoder_width=64, num_hidden_layers=1)
    w = qo.init_qformer_weights(cfg, seed=0, llm_dim=64)
    q, qt, proj = _build(cfg, w, 64)
    st = TrainableQFormer(q, qt, proj)
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(4, 16, 64, generator=g).cuda()
    ids = torch.randint(1000, 30000, (4, 8), generator=g).cuda()
    atts = torch.ones(4, 40, dtype=torch.long).cuda()
    G = torch.randn(4, 32, 64, generator=g).cuda()
    (st.forward(enc, ids, atts).float() * G).sum().backward()
    g1 = st.grad.clone()
    (st.forward(enc, ids, atts).float() * G).sum().backward()          # accumulates
    assert torch.allclose(st.grad, 2 * g1, rtol=1e-3, atol=1e-6 * g1.abs().max().item())
    st.zero_grad()
    (st.forward(enc, ids, atts).float() * (3 * G)).sum().backward()    # linear in the upstream gradient
    assert ((st.grad - 3 * g1).abs().max() / g1.abs().max()).item() < 2e-2
    # Adam step == torch.optim.Adam on the same flat buffers
    p0 = st.flat.clone()
    ref_p = torch.nn.Parameter(p0.clone())
    ref_p.grad = st.grad.clone()
    opt = torch.optim.Adam([ref_p], lr=3e-4)
    opt.step()
    st.adam_step(3e-4)
    assert (st.flat - ref_p.data).abs().max().item() < 1e-6
    assert torch.equal(st.flat16, st.flat.to(torch.bfloat16))
    # parameters alias the flat buffer: the module sees the update
    assert q.bert.encoder.layer[0].attention.self.query.weight.data_ptr() >= st.flat.data_ptr()


def test_trainer_step_reduces_surrogate_loss():
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                  llm_hidden_size=128, num_hidden_layers=2).cuda()
    tr = QFormerTrainer(model, accum_grad_iters=2, warmup_steps=0, init_lr=1e-3)
    g = torch.Generator().manual_seed(3)
    feats = {"video": torch.randn(2, 3, 17, 128, generator=g).cuda().to(torch.bfloat16),
             "audio": torch.randn(2, 3, 16, 64, generator=g).cuda().to(torch.bfloat16)}
    ids = torch.randint(1000, 30000, (2, 8), generator=g).cuda()
    mask = torch.ones(2, 8, dtype=torch.long).cuda()
    sur = {m: torch.randn(2, 3 * 32, 128, generator=g).cuda() for m in feats}
    losses = [tr.train_step(feats, ids, mask, surrogate=sur).item() for _ in range(8)]
    assert losses[-1] < losses[0], losses          # minimising sum(y * G): the loss must go down
    sd = tr.state_dict_trainable()
    assert "video_query_tokens" in sd and any(k.startswith("audio_Qformer.bert.encoder.layer.0") for k in sd)
    assert not any(k.endswith("_ln.weight") for k in sd)   # modality LNs stay frozen (models/xinstructblip.py:198-199)


def test_checkpoint_resume_and_eval_epoch(tmp_path):
    """Trainer shell: save -> load restores parameters, Adam moments and step count, so the next step
    continues identically (up to the rounding order of the atomics); eval_epoch parses generations with the reference's parser and scores them (utils/trainer.py:156-260)."""
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    from mraudio_b200 import mr_eval

    def make():
        torch.manual_seed(0)
        model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                      llm_hidden_size=128, num_hidden_layers=2).cuda()
        return QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-3)
    g = torch.Generator().manual_seed(3)
    feats = {"video": torch.randn(2, 2, 17, 128, generator=g).to(torch.bfloat16).cuda(),
             "audio": torch.randn(2, 2, 16, 64, generator=g).to(torch.bfloat16).cuda()}
    ids = torch.randint(1000, 30000, (2, 8), generator=g).cuda()
    mask = torch.ones(2, 8, dtype=torch.long).cuda()
    sur = {m: torch.randn(2, 2 * 32, 128, generator=g).cuda() for m in feats}
    a = make()
    for _ in range(2):
        a.train_step(feats, ids, mask, surrogate=sur)
    path = str(tmp_path / "ckpt" / "checkpoint_1.pth")
    a.save_checkpoint(path, cur_epoch=1)
    b = make()
    assert b.load_checkpoint(path) == 2
    for m in a.states:
        assert torch.equal(a.states[m].flat, b.states[m].flat) and b.states[m].step_count == 2
    la = a.train_step(feats, ids, mask, surrogate=sur)
    lb = b.train_step(feats, ids, mask, surrogate=sur)
    torch.cuda.synchronize()
    assert torch.equal(la, lb)
    for m in a.states:   # (fp32 atomics in the gradient reductions make two runs differ in the last bits)
        assert torch.allclose(a.states[m].flat, b.states[m].flat, rtol=0, atol=1e-5)
    gens = [(1, "q", "v", "[[10, 20]]", "[[10 20]]</s>"), (2, "q", "v", "[[0, 8], [30, 40]]", "[[30, 38]]"), (3, "q", "v", "[[4, 6]]", "nonsense")]
    res = a.eval_epoch(gens)
    sub = [{"qid": q, "pred_relevant_windows": p, "relevant_windows": t} for q, p, t in
           [(1, [[10, 20]], [[10, 20]]), (2, [[30, 38]], [[0, 8], [30, 40]]), (3, [[-1, -1]], [[4, 6]])]]
    assert res["brief"] == mr_eval.eval_submission(sub, sub, verbose=False)["brief"]
    assert res["brief"]["MR-full-invalid_pred_num"] == 1


def _small_trainer(accum=1, lr=1e-3):
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                  llm_hidden_size=128, num_hidden_layers=2).cuda()
    return QFormerTrainer(model, accum_grad_iters=accum, warmup_steps=0, init_lr=lr)


def _small_batch(seed=3):
    g = torch.Generator().manual_seed(seed)
    feats = {"video": torch.randn(2, 2, 17, 128, generator=g).to(torch.bfloat16).cuda(),
             "audio": torch.randn(2, 2, 16, 64, generator=g).to(torch.bfloat16).cuda()}
    ids = torch.randint(1000, 30000, (2, 8), generator=g).cuda()
    mask = torch.ones(2, 8, dtype=torch.long).cuda()
    sur = {m: torch.randn(2, 2 * 32, 128, generator=g).cuda() for m in feats}
    return feats, ids, mask, sur


def test_train_then_validate_then_train_uses_live_weights():
    """Validation between epochs (utils/trainer.py train_epoch / eval_epoch loop): the inference path must (a) see the
    weights the optimizer has just written through raw pointers and (b) not re-point the training handle at an inference
    snapshot.  Trainer A validates after every step, trainer B never does: identical losses, and every validation
    output equals the training forward of the same weights."""
    feats, ids, mask, sur = _small_batch()
    a, b = _small_trainer(), _small_trainer()
    outs = []
    for it in range(4):
        la = a.train_step(feats, ids, mask, surrogate=sur)
        lb = b.train_step(feats, ids, mask, surrogate=sur)
        # (from the third step on two identical trainers differ by the order of the gradient reduce-adds: see the note in
        #  test_overlapped_optimizer_matches_one_pass_training)
        assert torch.allclose(la, lb, rtol=1e-4 if it < 2 else 1e-3), (it, la.item(), lb.item())
        with torch.no_grad():
            inf, _ = a.model.encode_modalities(feats, ids, mask)
            trn, _ = a.forward_modalities(feats, ids, mask)
        for m in inf:
            # same bf16 weights, inference (fused LayerNorm, split residual) vs training (saved activations) forward
            assert ((inf[m].float() - trn[m].float()).abs().max() / trn[m].float().abs().max()).item() < 2e-2, (it, m)
        outs.append({m: inf[m].float().clone() for m in inf})
        a.states["video"]._saved = a.states["audio"]._saved = None
    # the validation outputs follow the optimizer: they change from step to step
    assert not torch.equal(outs[0]["video"], outs[-1]["video"])
    # the two-call projection form reads the live projection weights as well
    with torch.no_grad():
        h = torch.randn(3, 32, 768, device="cuda")
        y = a.model.video_llm_proj(h).float()
        ref = h.to(torch.bfloat16).float() @ a.model.video_llm_proj.weight.to(torch.bfloat16).float().t() + a.model.video_llm_proj.bias
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 1e-2
    # (the parameters of the two trainers are not compared element-wise: fp32 atomics make two runs differ in the last
    #  bits of a gradient, and Adam's normalisation turns a sign flip of a near-zero gradient into a full lr-sized step)


def test_checkpoint_optimizer_is_a_torch_adam_state_dict(tmp_path):
    """checkpoint['optimizer'] has torch.optim.Adam's state_dict layout (what utils/trainer.py:199,255 saves / loads):
    a stock Adam over the same module loads it, and a state_dict written by a stock Adam resumes here."""
    feats, ids, mask, sur = _small_batch()
    a = _small_trainer()
    for _ in range(3):
        a.train_step(feats, ids, mask, surrogate=sur)
    path = str(tmp_path / "c.pth")
    a.save_checkpoint(path, cur_epoch=0)
    ck = torch.load(path)
    osd = ck["optimizer"]
    assert set(osd) >= {"state", "param_groups"} and osd["param_groups"][0]["lr"] == a.lr
    params = list(a.model.parameters())
    opt = torch.optim.Adam(params, lr=3e-4)
    opt.load_state_dict({"state": osd["state"], "param_groups": osd["param_groups"]})      # stock Adam accepts it
    name_to_i = {n: i for i, n in enumerate(osd["param_names"])}
    i = name_to_i["video_Qformer.bert.encoder.layer.1.output_query.dense.weight"]
    st = opt.state[params[i]]
    assert float(st["step"]) == 3 and st["exp_avg"].shape == params[i].shape and st["exp_avg"].abs().max().item() > 0
    # frozen parameters (the modality LayerNorms) carry no state
    assert name_to_i["video_ln.weight"] not in osd["state"]
    # the reverse direction: the stock optimizer's state_dict resumes here, moments and step counts intact
    b = _small_trainer()
    b.load_optimizer_state_dict(opt.state_dict())
    for m in a.states:
        assert b.states[m].step_count == 3
        assert torch.equal(a.states[m].exp_avg, b.states[m].exp_avg) and torch.equal(a.states[m].exp_avg_sq, b.states[m].exp_avg_sq)
    # a round-1 style dict is rejected with a clear message
    with pytest.raises(ValueError, match="state_dict"):
        b.load_optimizer_state_dict({"video": {"exp_avg": None}})


def test_bucketed_adam_equals_one_pass():
    """The optimizer applied bucket by bucket behind the backward (TrainableQFormer.adam_bucket) is the one-pass
    adam_step: the buckets tile the flat buffer exactly once and every element gets the same arithmetic (bit-equal
    master weights, moments, bf16 operands, zeroed gradients) -- with the fp32 gradients and with bf16-exchanged ones."""
    a, b = _small_trainer(), _small_trainer()
    for m in a.states:
        sa, sb = a.states[m], b.states[m]
        cover = torch.zeros(sa.numel, dtype=torch.int32)
        for _, ranges in sa.buckets:
            for lo, hi in ranges:
                assert 0 <= lo < hi <= sa.numel and lo % 4 == 0 and hi % 4 == 0
                cover[lo:hi] += 1
        assert bool((cover == 1).all())
        for step, bf16 in enumerate((False, True, False)):
            g = torch.randn(sa.numel, device="cuda") * 1e-2
            sa.grad.copy_(g); sb.grad.copy_(g)
            if bf16:
                for s in (sa, sb):
                    s.grad16 = g.to(torch.bfloat16)
                sa._reduced_bf16 = True
            sa.adam_step(1e-3, grad_scale=0.5, zero_grad=True)
            sb.begin_bucketed_step()
            hyper = None
            if step == 2:   # the CUDA-graph form: lr / bias corrections / gradient scale read from device memory
                import ctypes
                from mraudio_b200._lib import lib
                host = torch.zeros(4)
                lib.mra_adam_hyper(1e-3, 0.9, 0.999, sb.step_count, 0.5, ctypes.cast(host.data_ptr(), ctypes.POINTER(ctypes.c_float)))
                hyper = host.cuda()
            for k in range(len(sb.buckets)):
                sb.adam_bucket(k, 1e-3 if hyper is None else 0.0, grad_scale=0.5 if hyper is None else 1.0, zero_grad=True,
                               reduced_bf16=bf16, hyper_dev=None if hyper is None else hyper.data_ptr())
            torch.cuda.synchronize()
            assert sa.step_count == sb.step_count == step + 1
            for name in ("flat", "exp_avg", "exp_avg_sq", "flat16", "grad"):
                assert torch.equal(getattr(sa, name), getattr(sb, name)), (m, step, name)
            assert float(sb.grad.abs().max()) == 0.0


def test_overlapped_optimizer_matches_one_pass_training():
    """train_step with the per-bucket optimizer on the communication stream against the one-pass optimizer after the
    backward: same losses step by step (the streams are ordered by the layer events; a missing dependency shows up as a
    diverging loss) and parameters equal up to the rounding order of the gradient atomics."""
    feats, ids, mask, sur = _small_batch()
    a, b = _small_trainer(lr=1e-4), _small_trainer(lr=1e-4)
    a.overlap_optimizer, b.overlap_optimizer = True, False
    # Tolerance = the run-to-run noise of ONE path: the weight gradients are summed by TMA reduce-add (split-K), whose order is
    # not fixed, so two identical trainers drift apart from the third step on (measured over 6 x 2 runs on a B200: steps 0 / 1
    # bit-equal, later steps up to 9.6e-5 relative for the same path twice, up to 1.2e-4 overlapped vs one-pass).  A missing
    # stream dependency lets Adam read half-written gradients: orders of magnitude above this.
    for it in range(5):
        la = a.train_step(feats, ids, mask, surrogate=sur)
        lb = b.train_step(feats, ids, mask, surrogate=sur)
        assert torch.allclose(la, lb, rtol=1e-6 if it < 2 else 1e-3), (it, la.item(), lb.item())
    torch.cuda.synchronize()
    for m in a.states:
        assert a.states[m].step_count == b.states[m].step_count == 5
        assert float(a.states[m].grad.abs().max()) == 0.0


@pytest.mark.parametrize("accum", [1, 2])
def test_cuda_graph_step_matches_eager(accum):
    """train_step replayed from CUDA graphs (forward | backward + per-bucket Adam, hyper-parameters in device memory)
    against the eager step: same losses while the warm-up schedule changes the learning rate every iteration and, with
    accum = 2, stepping and non-stepping iterations alternate; the step counts and the zeroed gradients agree."""
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers

    def make(graph):
        torch.manual_seed(0)
        model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                      llm_hidden_size=128, num_hidden_layers=2).cuda()
        return QFormerTrainer(model, accum_grad_iters=accum, warmup_steps=4, init_lr=2e-4, cuda_graph=graph)
    a, b = make(True), make(False)
    batches = [_small_batch(seed) for seed in (3, 4, 5)]
    for it in range(8):
        feats, ids, mask, sur = batches[it % 3]
        la = a.train_step(feats, ids, mask, surrogate=sur)
        lb = b.train_step(feats, ids, mask, surrogate=sur)
        assert a.lr == b.lr
        # (two eager runs differ by ~1e-4 after a few steps: fp32 atomics in the gradient reductions + Adam's normalisation)
        assert torch.allclose(la, lb, rtol=1e-3), (it, la.item(), lb.item())
    torch.cuda.synchronize()
    assert len(a._graphs) == 1                       # one shape -> one pair of graphs, reused for every batch
    for m in a.states:
        assert a.states[m].step_count == b.states[m].step_count == 8 // accum
        assert float(a.states[m].grad.abs().max()) == 0.0
        # the second moments are a smooth function of the gradient history: they agree closely if every replay saw the
        # right inputs, gradients and bias corrections
        va, vb = a.states[m].exp_avg_sq, b.states[m].exp_avg_sq
        assert ((va - vb).abs().max() / vb.abs().max()).item() < 2e-2, m
        rel = ((a.states[m].flat - b.states[m].flat).abs().max() / b.states[m].flat.abs().max()).item()
        assert rel < 5e-3, (m, rel)
    # validation through the inference path sees the graph-updated weights
    feats, ids, mask, sur = batches[0]
    with torch.no_grad():
        ia, _ = a.model.encode_modalities(feats, ids, mask)
        ib, _ = b.model.encode_modalities(feats, ids, mask)
    for m in ia:
        assert ((ia[m].float() - ib[m].float()).abs().max() / ib[m].float().abs().max()).item() < 3e-2
