"""Secondary measurements printed by bench.py under the "secondary" key of its JSON line (same torchrun world, after the
headline): the paths of BASELINE.json's configs 3, 4 and 5 that the headline forward does not touch -- in particular the
two places where the path has a real exchange step: the fine-tuning gradient all-reduce and the gather of
scored moments.  Every time is a CUDA-event time on the launching stream, max over ranks, after warm-up.

  cfg4  fine-tuning step (utils/trainer.py:124-140): 8 videos x 8 frames per GPU, both modalities, forward + backward +
        gradient all-reduce + Adam; the same step with the all-reduce skipped (exposed all-reduce = difference); NCCL bytes
        per step; at N > 1 the data-parallel equivalence check: the all-reduced gradient of the ranks' shards == the
        gradient of the whole batch computed on one rank (the defining property of DistributedDataParallel, :69)
  cfg5  moment-retrieval sweep: 16 videos x 75 clips per GPU per step through both Q-Formers + projections, then parse of the
        (synthetic) generations, GPU R1 / mAP scoring and the fixed-width gather of the records to rank 0 (:163-181)
  cfg3  Video-LLaMA-v1-style video (32 frames, frame position embedding) + ImageBind-audio Q-Formers at batch 64 per GPU
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Ctx:
    def __init__(self, world, rank, dev):
        self.world, self.rank, self.dev = world, rank, dev

    def barrier(self):
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) / steps)


def _train_batch(B, F, T, seed, dev, scale=1e-3, text_seed=None):
    g = torch.Generator().manual_seed(seed)
    feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
             "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
    if text_seed is None:
        ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
    else:   # the same instruction prompt for every video (of every rank)
        ids = torch.randint(1000, 30000, (1, T), generator=torch.Generator().manual_seed(text_seed)).expand(B, T).contiguous().to(dev)
    mask = torch.ones(B, T, dtype=torch.long, device=dev)
    sur = {m: (torch.randn(B, F * 32, 4096, generator=g) * scale).to(dev) for m in feats}
    return feats, ids, mask, sur


def cfg4_finetune(cx: Ctx, steps=8, warmup=3):
    from mraudio_b200.training import QFormerTrainer
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    B, F, T = 8, 8, 32
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio")).to(cx.dev)
    tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-5, grad_comm_dtype=torch.bfloat16)
    feats, ids, mask, sur = _train_batch(B, F, T, 1 + cx.rank, cx.dev)
    step = lambda: tr.train_step(feats, ids, mask, surrogate=sur)
    out = {"workload": "finetune.py step on cached features: 8 videos x 8 frames per GPU, video + audio Q-Former + llm_proj, "
                       "fwd + bwd + gradient all-reduce + Adam (372 M parameters), surrogate loss (LLM out of scope)",
           "n_gpus": cx.world}
    if cx.world > 1:
        # ---- data-parallel equivalence on a small batch, BEFORE anything lets the replicas drift: mean over ranks of the shard
        #      gradients == gradient of the whole batch (fp32 exchange for this check; with the sum-loss the whole-batch
        #      gradient is the SUM of the shard gradients).  Every video carries the same prompt: the reference tiles the
        #      prompt frame-major over the (video, frame) rows (models/xinstructblip.py:287-289), so with DIFFERENT prompts
        #      per video the pairing of prompt and row depends on the batch size, and a shard would not compute the same
        #      function as the whole batch.
        Bc = 2
        tr.set_grad_comm_dtype(torch.float32)
        parts = [_train_batch(Bc, F, T, 1000 + r, cx.dev, text_seed=77) for r in range(cx.world)]
        shard = parts[cx.rank]
        tr.train_step(shard[0], shard[1], shard[2], surrogate=shard[3], apply_optimizer=False)
        torch.cuda.synchronize()
        g_ddp = {m: st.grad.clone() for m, st in tr.states.items()}
        whole = ({m: torch.cat([p[0][m] for p in parts]) for m in parts[0][0]}, torch.cat([p[1] for p in parts]),
                 torch.cat([p[2] for p in parts]), {m: torch.cat([p[3][m] for p in parts]) for m in parts[0][3]})
        for st in tr.states.values():
            st.zero_grad()
        tr.allreduce_enabled = False
        tr.train_step(whole[0], whole[1], whole[2], surrogate=whole[3], apply_optimizer=False)
        tr.allreduce_enabled = True
        torch.cuda.synchronize()
        worst = 0.0
        for m, st in tr.states.items():
            worst = max(worst, ((g_ddp[m] - st.grad).abs().max() / st.grad.abs().max()).item())
        worst = cx.max_over_ranks(worst)
        # (recorded, not raised: a failed check must not take the headline line down; tests/test_gpu_multi.py asserts it)
        out["ddp_equivalence"] = {"check": f"all-reduced gradient of {cx.world} shards of {Bc} videos == gradient of the {cx.world * Bc}-video batch "
                                           "computed on one rank, all 186 M parameters of each modality, fp32 exchange, max-norm relative",
                                  "max_rel_err": worst, "tolerance": 1e-5, "ok": bool(worst < 1e-5)}
        for st in tr.states.values():
            st.zero_grad()
        del g_ddp, parts, whole, shard
        tr.set_grad_comm_dtype(torch.bfloat16)
    ms = cx.timed(step, steps, warmup)
    out["ms_per_step"] = ms
    out["clips_per_s_all_gpus"] = cx.world * B * F / (ms * 1e-3)
    out["backward_launches"] = sum(s.last_backward_launches for s in tr.states.values())
    numel = sum(s.numel for s in tr.states.values())
    out["grad_allreduce"] = {"dtype": "bf16 (fp32 master weights / Adam state)", "buckets_per_modality": len(next(iter(tr.states.values())).buckets),
                             "nccl_bytes_per_step": (numel * 2) if cx.world > 1 else 0}
    if cx.world > 1:
        tr.overlap_allreduce = False
        out["ms_per_step_flat_allreduce_after_backward"] = cx.timed(step, steps, 2)
        tr.overlap_allreduce = True
        tr.allreduce_enabled = False      # last: without the exchange the replicas drift apart
        ms_no = cx.timed(step, steps, 2)
        tr.allreduce_enabled = True
        out["ms_per_step_without_allreduce"] = ms_no
        out["exposed_allreduce_ms"] = ms - ms_no
        out["scaling_vs_no_exchange"] = ms_no / ms
    else:
        # single process: the same step replayed from CUDA graphs (forward | eager loss | backward + per-bucket Adam)
        tr.cuda_graph = True
        out["ms_per_step_cuda_graph_replay"] = cx.timed(step, steps, 3)
        tr.cuda_graph = False
    del tr, model
    torch.cuda.empty_cache()
    return out


def _synth_generation(rng, n_max):
    """an LLM-style generation: 1..n_max windows on the 2-s grid of a 150-s video, with the occasional format slip"""
    wins = []
    for _ in range(int(rng.integers(1, n_max + 1))):
        a, b = sorted(rng.integers(0, 76, size=2).tolist())
        b = max(b, a + 1)
        wins.append(f"[{2 * a} {2 * b}]" if rng.random() < 0.1 else f"[{2 * a}, {2 * b}]")
    return "junk" if rng.random() < 0.02 else "[" + ", ".join(wins) + "]</s>"


def cfg5_sweep(cx: Ctx, steps=3, warmup=1, queries_per_video=100):
    from mraudio_b200 import mr_eval
    from mraudio_b200.parsing import parse_output
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers
    B, F, T = 16, 75, 32
    torch.manual_seed(0)
    model = XInstructBLIPQFormers(modalities=("video", "audio")).to(cx.dev).eval()
    g = torch.Generator().manual_seed(100 + cx.rank)
    feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(cx.dev),
             "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(cx.dev)}
    ids = torch.randint(1000, 30000, (B, T), generator=g).to(cx.dev)
    mask = torch.ones(B, T, dtype=torch.long, device=cx.dev)

    def fwd():
        with torch.no_grad():
            model.encode_modalities(feats, ids, mask)
    ms_fwd = cx.timed(fwd, steps, warmup)
    rng = np.random.default_rng(7 + cx.rank)
    n_local = B * queries_per_video
    base = cx.rank * n_local
    t0 = time.perf_counter()
    records = [{"qid": base + i, "_order": base + i, "pred_relevant_windows": parse_output(_synth_generation(rng, 5)),
                "relevant_windows": parse_output(_synth_generation(rng, 3).replace("junk", "[[0, 2]]"))} for i in range(n_local)]
    t_parse = time.perf_counter() - t0
    mr_eval.score_records_distributed(records[:64], records[:64])     # warm-up: gather channels, pinned staging
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    rec = mr_eval.score_records_distributed(records, records)
    e1.record()
    torch.cuda.synchronize()
    t_wall = cx.max_over_ranks(time.perf_counter() - t0)
    ms_dev = cx.max_over_ranks(e0.elapsed_time(e1))
    out = {"workload": "moment-retrieval sweep: 150-s videos at 0.5 fps (75 clips), 16 videos per GPU per step through both Q-Formers + "
                       "projections; then parse + GPU R1 / mAP scoring + gather of the records to rank 0",
           "n_gpus": cx.world, "forward_ms_per_step": ms_fwd, "clips_per_s_all_gpus": cx.world * B * F / (ms_fwd * 1e-3),
           "videos_per_s_all_gpus": cx.world * B / (ms_fwd * 1e-3), "scored_queries_all_gpus": n_local * cx.world,
           "parse_us_per_generation": t_parse / (2 * n_local) * 1e6,
           "score_and_gather_ms": {"device_events": ms_dev, "host_wall_incl_packing": t_wall * 1e3},
           "gather_bytes_to_rank0": n_local * (cx.world - 1) * 13 * 8}
    if cx.rank == 0:
        total = n_local * cx.world
        stub = [{"qid": i, "pred_relevant_windows": [[0, 0]], "relevant_windows": [[0, 0]]} for i in range(total)]
        res = mr_eval.eval_submission(stub, stub, verbose=False, _records=rec)
        out["brief"] = {k: res["brief"][k] for k in ("MR-full-R1@0.5", "MR-full-R1@0.7", "MR-full-mAP", "MR-full-invalid_pred_num")}
    del model, feats
    torch.cuda.empty_cache()
    return out


def cfg3_videollama(cx: Ctx, steps=10, warmup=3):
    from mraudio_b200.videollama import VideoLLaMAQFormers
    B, F = 64, 32
    torch.manual_seed(0)
    vl = VideoLLaMAQFormers().to(cx.dev).eval()
    g = torch.Generator().manual_seed(300 + cx.rank)
    frames = torch.randn(B, F, 32, 768, generator=g).to(torch.bfloat16).to(cx.dev)
    audio = torch.randn(B, 8, 1024, generator=g).to(torch.bfloat16).to(cx.dev)

    def step():
        with torch.no_grad():
            vl.encode_videoQformer(frames)
            vl.encode_audioQformer(audio)
    ms = cx.timed(step, steps, warmup)
    del vl
    torch.cuda.empty_cache()
    return {"workload": "Video-LLaMA-v1-style video Q-Former (32 frames x 32 tokens + frame position embedding -> 1024 keys, 2 layers) + "
                        "ImageBind-audio Q-Former, batch 64 per GPU, bf16 (parity unpinned w.r.t. the reference, see DESIGN.md)",
            "n_gpus": cx.world, "ms_per_step": ms, "videos_per_s_all_gpus": cx.world * B / (ms * 1e-3)}


def run_all(world, rank, dev):
    cx = Ctx(world, rank, dev)
    out = {}
    for name, fn in (("cfg4_finetune_step", cfg4_finetune), ("cfg5_eval_sweep", cfg5_sweep), ("cfg3_videollama_v1", cfg3_videollama)):
        try:
            out[name] = fn(cx)
        except Exception as e:   # a secondary measurement must not take the headline down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    return out
