"""Secondary measurements for DESIGN.md (not the headline): fine-tuning step (config 4 shape), the config-5 sweep shape,
config 3 (Video-LLaMA-v1-style) and the scorer.  Run on the GPU box:  python tools/extra_bench.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mraudio_b200 import mr_eval, ops
from mraudio_b200.training import QFormerTrainer
from mraudio_b200.videollama import VideoLLaMAQFormers
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

dev = torch.device("cuda:0")
out = {}


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- config 4: fine-tuning step on cached features, per-GPU batch 8 videos x 8 frames, T = 32, surrogate loss
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev)
tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0)
g = torch.Generator().manual_seed(1)
B, F, T = 8, 8, 32
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
mask = torch.ones(B, T, dtype=torch.long, device=dev)
sur = {m: torch.randn(B, F * 32, 4096, generator=g).to(dev) for m in feats}
ms = timeit(lambda: tr.train_step(feats, ids, mask, surrogate=sur), n=5, warm=2)
fwd_flop = B * F * 33.95e9
out["cfg4_train_step"] = {"ms_per_step": ms, "clips_per_s": B * F / (ms * 1e-3), "videos_per_gpu": B, "frames": F,
                           "approx_tflops": 3 * fwd_flop / (ms * 1e-3) / 1e12,
                           "backward_launches": sum(s.last_backward_launches for s in tr.states.values())}
del tr, model, feats, sur
torch.cuda.empty_cache()

# ---- config 5 shape: 16 videos x 75 clips per GPU = 1200 rows per modality
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev).eval()
B, F = 16, 75
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
mask = torch.ones(B, T, dtype=torch.long, device=dev)
with torch.no_grad():
    ms = timeit(lambda: model.encode_modalities(feats, ids, mask), n=3, warm=1)
out["cfg5_sweep_shape"] = {"ms_per_step": ms, "clips_per_s": B * F / (ms * 1e-3), "videos_per_gpu": B, "clips_per_video": F,
                            "frac_of_sustained_peak": B * F * 33.95e9 / (ms * 1e-3) / 1367.2e12}
del model, feats
torch.cuda.empty_cache()

# ---- config 3: Video-LLaMA-v1-style, batch 64, 32 frames
vl = VideoLLaMAQFormers().to(dev).eval()
frames = torch.randn(64, 32, 32, 768, generator=g).to(torch.bfloat16).to(dev)
audio = torch.randn(64, 8, 1024, generator=g).to(torch.bfloat16).to(dev)
with torch.no_grad():
    ms = timeit(lambda: (vl.encode_videoQformer(frames), vl.encode_audioQformer(audio)), n=10, warm=3)
out["cfg3_videollama_v1"] = {"ms_per_step": ms, "videos_per_s": 64 / (ms * 1e-3)}

# ---- scorer: config-5 sized sweep (128 videos x 400 queries)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mr_eval_oracle as mo
Q = 51200
sub, gt = mo.synth_submission(Q, seed=7)
t0 = time.perf_counter(); rec = mr_eval.score_records(sub, gt); t_api = time.perf_counter() - t0
pred, npred, gtw, ngt = mr_eval.pack_windows(sub, gt)
dv = [torch.from_numpy(a).to(dev) for a in (pred, npred, gtw, ngt)]
thd = torch.tensor(mr_eval.IOU_THDS, dtype=torch.float64, device=dev)
ms_k = timeit(lambda: ops.mr_score(dv[0], dv[1], dv[2], dv[3], thd), n=20, warm=3)
n_cpu = 2000
t0 = time.perf_counter(); mo.score_records(sub[:n_cpu], gt[:n_cpu]); t_cpu = time.perf_counter() - t0
bytes_q = pred[0].nbytes + gtw[0].nbytes + 8 + 89
out["scorer"] = {"queries": Q, "kernel_ms": ms_k, "kernel_queries_per_s": Q / (ms_k * 1e-3), "kernel_gbs": Q * bytes_q / (ms_k * 1e-3) / 1e9,
                 "api_s_incl_host_packing": t_api, "cpu_oracle_queries_per_s": n_cpu / t_cpu, "cpu_sample_queries": n_cpu}
print(json.dumps(out, indent=1))
