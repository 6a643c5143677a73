"""Micro-benchmark of the attention kernels on the shapes of the Q-Former step (tuning aid; run on the GPU box): row-major
operands ([tokens, heads * 64], the training forward) against head-major ones ([head][token][64], the inference forward).
Env: MRA_ATT_STAGES / MRA_ATT_OSTAGE / MRA_LIB select the experiments (see csrc/attention.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows, heads, H = int(os.environ.get("ROWS", "512")), 12, 768
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("MRA_"))


def timeit(fn):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[3] * 1e3


for name, Sk, ld in (("cross_video", 257, 9216), ("cross_audio", 256, 9216)):
    q = torch.randn(rows * 32, H, device=dev).to(torch.bfloat16)
    kv = torch.randn(rows * Sk, ld, device=dev).to(torch.bfloat16)
    byt = rows * Sk * 2 * H * 2 + 2 * q.numel() * 2
    t = timeit(lambda: ops.attention(q, kv[:, :H], kv[:, H:2 * H], rows, heads, 32, Sk, 32, True))
    print(f"[{tag}] {name} rows={rows} row-major  {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s")
    hm = kv.view(rows * Sk, ld // 64, 64).permute(1, 0, 2).contiguous()
    qv = q.view(rows * 32, heads, 64).permute(1, 0, 2)
    t = timeit(lambda: ops.attention_head_major(qv, hm[:heads], hm[heads:2 * heads], rows, 32, Sk, 32, True))
    print(f"[{tag}] {name} rows={rows} head-major {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s", flush=True)
    del kv, hm
qkv = torch.randn(rows * 64, 3 * H, device=dev).to(torch.bfloat16)
byt = qkv.numel() * 2 + rows * 64 * H * 2
t = timeit(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], rows, heads, 64, 64, 32, False))
print(f"[{tag}] self S=64 rows={rows} row-major  {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s")
hm = qkv.view(rows * 64, 3 * heads, 64).permute(1, 0, 2).contiguous()
t = timeit(lambda: ops.attention_head_major(hm[:heads], hm[heads:2 * heads], hm[2 * heads:], rows, 64, 64, 32, False))
print(f"[{tag}] self S=64 rows={rows} head-major {t:7.1f} us  {byt / t / 1e3:6.0f} GB/s", flush=True)
