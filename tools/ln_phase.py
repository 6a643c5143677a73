"""Per-phase cycle counts of the fused GEMM+LayerNorm kernel (MRA_LN_DEBUG=8 instrumentation: CTA 0, epilogue warp 0)."""
import os, sys
os.environ["MRA_LN_DEBUG"] = os.environ.get("MRA_LN_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
for name, M, K in [("ao_x2", 32768, 768), ("f2_x4", 32768, 3072), ("co_x2", 16384, 768)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(768, K, device=dev) * 0.03).bfloat16()
    b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev); g = torch.ones(768, device=dev); be = torch.zeros(768, device=dev)
    for _ in range(3):
        sys.stderr.write(f"{name}: "); sys.stderr.flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12); e1.record(); torch.cuda.synchronize()
    print(name, "last launch incl. sync overhead: %.1f us" % (e0.elapsed_time(e1) * 1e3))
