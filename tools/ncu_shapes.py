"""A handful of single launches of the N = 768 Linear shapes for an `ncu --set full` capture (not a bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
torch.manual_seed(0)
def mk(M, N, K):
    return (torch.randn(M, K, device=dev).bfloat16(), (torch.randn(N, K, device=dev) * 0.03).bfloat16(), torch.randn(N, device=dev),
            torch.randn(M, N, device=dev))
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name, M, N, K in [("f2x2", 32768, 768, 3072), ("ao_x2", 32768, 768, 768)]:
    x, w, b, r = mk(M, N, K)
    g, be = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    for it in range(2):
        big.zero_()                                   # flush L2
        ops.linear(x, w, b, residual=r, out_fp32=True)   # plain GEMM, fp32 out + residual epilogue
        big.zero_()
        ops.linear(x, w, b)                            # plain GEMM, bf16 out
        big.zero_()
        ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12)
    torch.cuda.synchronize()
print("done")
