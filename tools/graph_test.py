"""Does replaying the step as a CUDA graph change the step time?  (host launch overhead vs GPU-side gaps)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200.xinstructblip import XInstructBLIPQFormers
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev).eval()
g = torch.Generator().manual_seed(1)
feats = {"video": torch.randn(32, 8, 257, 1408, generator=g).to(torch.bfloat16).to(dev), "audio": torch.randn(32, 8, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (32, 32), generator=g).to(dev); mask = torch.ones(32, 32, dtype=torch.long, device=dev)
def step():
    with torch.no_grad():
        return model.encode_modalities(feats, ids, mask)
for _ in range(3): step()
torch.cuda.synchronize()
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager  ms/step", timeit(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2): step()
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out = step()
torch.cuda.synchronize()
print("graph  ms/step", timeit(graph.replay))
import time
t0 = time.perf_counter(); step(); t1 = time.perf_counter()
print("host enqueue time of one eager step: %.2f ms" % ((t1 - t0) * 1e3))
