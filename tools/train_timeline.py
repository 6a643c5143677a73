"""Device timeline (CUPTI via torch.profiler) of one fine-tuning step of one modality at config 4's per-GPU shape
(8 videos x 8 frames = 64 rows): per kernel type count / total / average duration and the idle gaps, forward, backward
and optimizer separately.  Not a bench number."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from mraudio_b200.training import TrainableQFormer
from mraudio_b200.xinstructblip import XInstructBLIPQFormers
dev = torch.device("cuda:0")
torch.manual_seed(0)
mod = os.environ.get("TL_MODALITY", "video")
Nk, W = (257, 1408) if mod == "video" else (256, 768)
model = XInstructBLIPQFormers(modalities=(mod,)).to(dev)
model.freeze_qformers(False)
st = TrainableQFormer(getattr(model, f"{mod}_Qformer"), getattr(model, f"{mod}_query_tokens"), getattr(model, f"{mod}_llm_proj"))
g = torch.Generator().manual_seed(1)
rows, T = int(os.environ.get("TL_ROWS", "64")), 32
enc = torch.randn(rows, Nk, W, generator=g).to(torch.bfloat16).to(dev)
ids = torch.randint(1000, 30000, (rows, T), generator=g).to(dev)
atts = torch.ones(rows, 32 + T, dtype=torch.long, device=dev)
G = torch.randn(rows, 32, 4096, generator=g).to(dev)


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def step(timed=False):
    e0 = ev(); y = st.forward(enc, ids, atts); e1 = ev()
    loss = (y.float() * G).sum(); e2 = ev()
    loss.backward(); e3 = ev()
    st.adam_step(1e-4, zero_grad=True); e4 = ev(); e5 = ev()
    if timed:
        torch.cuda.synchronize()
        print(f"fwd {e0.elapsed_time(e1):.2f} | loss {e1.elapsed_time(e2):.2f} | bwd {e2.elapsed_time(e3):.2f} | adam+refresh "
              f"{e3.elapsed_time(e4):.2f} | zero {e4.elapsed_time(e5):.2f} | total {e0.elapsed_time(e5):.2f} ms (bwd launches {st.last_backward_launches})")


for _ in range(3):
    step()
for _ in range(3):
    step(True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev_ = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev_.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
prev_end = ev_[0].time_range.start
t0 = prev_end
for e in ev_:
    d = e.time_range.end - e.time_range.start
    gap = max(0.0, e.time_range.start - prev_end)
    prev_end = max(prev_end, e.time_range.end)
    name = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("mra::", "").split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d; a[2] += gap
print(f"span {prev_end - t0:.0f} us, kernels {len(ev_)}, busy {sum(a[1] for a in agg.values()):.0f} us, gaps {sum(a[2] for a in agg.values()):.0f} us")
for k, (n, d, gp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={n:4d} dur={d:8.1f} avg={d / n:7.1f} gap_before_tot={gp:7.1f}")

seqs = collections.OrderedDict()
for e in ev_:
    n = e.name.replace("void ", "").replace("(anonymous namespace)::", "").replace("mra::", "").split("(")[0][:40]
    if "attn_bwd" in n or "ln_bwd" in n or "attention_tma" in n:
        seqs.setdefault(n, []).append(round(e.time_range.end - e.time_range.start, 1))
for n, l in seqs.items():
    print("per-launch us", n, l)
