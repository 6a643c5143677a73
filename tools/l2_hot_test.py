"""Does an L2-resident A operand make the narrow GEMMs faster?  (decides whether evict-last hints on the producers' bf16
outputs are worth having)  A is either evicted (a 256 MB memset right before the launch) or freshly written (copy)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, gelu in [("ffn_up", 32768, 3072, 768, True), ("qkv", 32768, 2304, 768, False), ("cross_q", 16384, 768, 768, False)]:
    x = torch.randn(M, K, device=dev).bfloat16(); x2 = x.clone()
    w = (torch.randn(N, K, device=dev) * 0.03).bfloat16(); b = torch.randn(N, device=dev)
    res = {}
    for mode in ("cold", "hot"):
        ts = []
        for it in range(7):
            flush.zero_()
            if mode == "hot":
                x.copy_(x2)          # A freshly written: resident in L2 (50 MB of 126 MB)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.linear(x, w, b, gelu=gelu); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[mode] = sorted(ts[2:])[2]
    print(f"{name:8s} M={M} N={N} K={K}: A cold {res['cold']:.1f} us, A hot in L2 {res['hot']:.1f} us")
