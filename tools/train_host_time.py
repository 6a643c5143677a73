"""Host enqueue time against device time of the config-4 fine-tuning step (is the step bound by the CPU issuing ~800 launches?)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200.training import QFormerTrainer
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev)
g = torch.Generator().manual_seed(1)
B, F, T = 8, 8, 32
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
mask = torch.ones(B, T, dtype=torch.long, device=dev)
sur = {m: (torch.randn(B, F * 32, 4096, generator=g) * 1e-3).to(dev) for m in feats}
tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-5)
for _ in range(3):
    tr.train_step(feats, ids, mask, surrogate=sur)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    tr.train_step(feats, ids, mask, surrogate=sur)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"host_enqueue_ms_per_step": (t1 - t0) * 1e3 / n, "device_ms_per_step": e0.elapsed_time(e1) / n,
                  "wall_ms_per_step": (t2 - t0) * 1e3 / n, "cpus": os.cpu_count()}))
