import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, K in (("f2_x4", 32768, 3072), ("ao_x2", 32768, 768)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(768, K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev); g = torch.ones(768, device=dev); be = torch.zeros(768, device=dev)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(os.environ.get("MRA_LN_DEBUG", "0"), name, f"{sorted(ts[1:])[2]*1e3:7.1f} us")
