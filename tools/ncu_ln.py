"""Single launches for an `ncu --set full` capture: fused GEMM+LayerNorm (split residual stream) on the FFN-down and the
attention-output shapes of a config-2 step, and the plain tcgen05 GEMM on the FFN-down shape and on a wide-N shape for
comparison.  L2 flushed before each launch.  Not a bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
torch.manual_seed(0)
big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
g, be = torch.ones(768, device=dev), torch.zeros(768, device=dev)
for rep in range(2):
    for name, M, K in [("f2_x4", 32768, 3072), ("ao_x2", 32768, 768)]:
        x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(768, K, device=dev) * 0.03).bfloat16()
        b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev)
        rh, rl = ops.split_residual(r)
        big.zero_(); ops.linear_residual_layernorm_split(x, w, b, rh, rl, g, be, 1e-12)
        big.zero_(); ops.linear(x, w, b)                                  # plain GEMM, same shape, bf16 out
    x = torch.randn(32768, 768, device=dev).bfloat16(); w = (torch.randn(3072, 768, device=dev) * 0.03).bfloat16()
    b = torch.randn(3072, device=dev)
    big.zero_(); ops.linear(x, w, b, gelu=True)                           # FFN-up with the GELU epilogue
    big.zero_(); ops.linear(x, w, b)                                      # the same without GELU
torch.cuda.synchronize()
print("done")
