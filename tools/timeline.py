"""Device timeline of one config-2 step (CUPTI activity records through torch.profiler): per kernel start / duration and the
idle gap before it, so that time between kernels (launch latency, tails) can be told apart from time inside kernels.
Not a bench number (the profiler adds overhead); used to decide what to fuse / overlap next."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, F = int(os.environ.get("TL_VIDEOS", "32")), int(os.environ.get("TL_FRAMES", "8"))
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev).eval()
g = torch.Generator().manual_seed(1)
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, 32), generator=g).to(dev)
mask = torch.ones(B, 32, dtype=torch.long, device=dev)


def step():
    with torch.no_grad():
        return model.encode_modalities(feats, ids, mask)


for _ in range(3):
    step()
torch.cuda.synchronize()
NSTEP = int(os.environ.get("TL_STEPS", "10"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NSTEP):
    step()
e1.record()
torch.cuda.synchronize()
print(f"unprofiled: {e0.elapsed_time(e1) / NSTEP:.3f} ms/step over {NSTEP} steps")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "mem" not in e.name.lower()[:6]]
ev.sort(key=lambda e: e.time_range.start)
# last step only: kernels after the last embed_ln pair
starts = [i for i, e in enumerate(ev) if "embed_ln" in e.name][::2]
for a, b in zip(starts, starts[1:] + [len(ev)]):
    print(f"step span {ev[b - 1].time_range.end - ev[a].time_range.start:.1f} us, busy {sum(e.time_range.end - e.time_range.start for e in ev[a:b]):.1f} us")
ev = ev[starts[-1]:]
t0 = ev[0].time_range.start
agg = collections.OrderedDict()
rows = []
prev_end = t0
for e in ev:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = s - prev_end
    prev_end = max(prev_end, e.time_range.end)
    name = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("mra::", "").split("(")[0][:60]
    rows.append((round(s - t0, 1), round(d, 1), round(gap, 1), name))
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d
    a[2] += gap
span = prev_end - t0
print(f"step span {span:.1f} us, kernels {len(ev)}, busy {sum(r[1] for r in rows):.1f} us, gaps {sum(r[2] for r in rows):.1f} us")
for k, (n, d, gp) in agg.items():
    print(f"{k:62s} n={n:3d} dur={d:8.1f} avg={d / n:7.1f}  gap_before avg={gp / n:5.1f} tot={gp:7.1f}")
if os.environ.get("TL_DUMP"):
    for r in rows:
        print(*r)
