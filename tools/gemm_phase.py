"""Per-tile cycles of the plain GEMM's epilogue warps: blocked on the accumulator vs working (MRA_GEMM_DEBUG=8)."""
import os, sys
os.environ.setdefault("MRA_GEMM_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops
dev = "cuda"
for name, M, N, K, gelu in [("ffn_up", 32768, 3072, 768, True), ("qkv", 32768, 2304, 768, False), ("kv_audio", 65536, 9216, 768, False),
                            ("kv_video", 65792, 9216, 1408, False), ("proj", 16384, 4096, 768, False)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.03).bfloat16(); b = torch.randn(N, device=dev)
    for _ in range(3):
        sys.stderr.write(name + " "); sys.stderr.flush()
        ops.linear(x, w, None if os.environ.get("GP_NOBIAS") else b, gelu=gelu)
    torch.cuda.synchronize()
