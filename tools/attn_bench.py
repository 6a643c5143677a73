"""Micro-benchmark of the attention kernels on the shapes of the Q-Former step (tuning aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows, heads, H = 256, 12, 768


def timeit(fn):
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[2] * 1e3


for name, Sk, ld in (("cross_video", 257, 9216), ("cross_audio", 256, 9216)):
    q = torch.randn(rows * 32, H, device=dev).to(torch.bfloat16)
    kv = torch.randn(rows * Sk, ld, device=dev).to(torch.bfloat16)
    byt = rows * Sk * 2 * H * 2 + 2 * q.numel() * 2
    for impl in (0, 1):
        _lib.lib.mra_attention_impl_override(impl)
        t = timeit(lambda: ops.attention(q, kv[:, :H], kv[:, H:2 * H], rows, heads, 32, Sk, 32, True))
        print(f"{name} impl={'generic' if impl else 'tma'} {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s")
qkv = torch.randn(rows * 64, 3 * H, device=dev).to(torch.bfloat16)
byt = qkv.numel() * 2 + rows * 64 * H * 2
for impl in (0, 1):
    _lib.lib.mra_attention_impl_override(impl)
    t = timeit(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], rows, heads, 64, 64, 32, False))
    print(f"self S=64 impl={'generic' if impl else 'tma'} {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s")
_lib.lib.mra_attention_impl_override(0)
