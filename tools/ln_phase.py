"""Per-phase cycle counts of the fused GEMM+LayerNorm kernel (instrumented build: MRA_LIB=instr, MRA_LN_DEBUG=8; CTA 0,
epilogue warp 0), fp32 and split residual streams."""
import os, sys
os.environ.setdefault("MRA_LIB", "instr")
os.environ["MRA_LN_DEBUG"] = os.environ.get("MRA_LN_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
for name, M, K in [("ao_x2", 32768, 768), ("f2_x4", 32768, 3072), ("co_x2", 16384, 768)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(768, K, device=dev) * 0.03).bfloat16()
    b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev); g = torch.ones(768, device=dev); be = torch.zeros(768, device=dev)
    rh, rl = ops.split_residual(r)
    for form in ("fp32", "split"):
        for _ in range(2):
            sys.stderr.write(f"{name} {form}: "); sys.stderr.flush()
            if form == "fp32":
                ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12)
            else:
                ops.linear_residual_layernorm_split(x, w, b, rh, rl, g, be, 1e-12)
            torch.cuda.synchronize()
