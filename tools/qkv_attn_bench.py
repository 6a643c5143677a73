"""Fused QKV Linear + self-attention core (csrc/qkv_attn.cu) against the two launches it replaces, on the shape of a config-2
step (512 rows = both modalities, 64 tokens per row, hidden 768); L2 flushed before every timed region (tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows, heads, H = int(os.environ.get("ROWS", "512")), 12, 768


def timeit(fn):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[3] * 1e3


x = torch.randn(rows * 64, H, device=dev).to(torch.bfloat16)
w = (torch.randn(3 * H, H, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(3 * H, device=dev)
mask = torch.zeros(rows, 64, device=dev)
fl = 2 * rows * 64 * 3 * H * H


def unfused():
    qkv = ops.linear(x, w, b)
    return ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], rows, heads, 64, 64, 32, False, mask)


def unfused_hm():
    qkv = ops.linear_head_major(x, w, b)
    return ops.attention_head_major(qkv[:heads], qkv[heads:2 * heads], qkv[2 * heads:], rows, 64, 64, 32, False, mask)


a, c = unfused(), ops.qkv_self_attention(x, w, b, rows, mask)
print("bit-equal:", bool(torch.equal(a, c)))
t_lin = timeit(lambda: ops.linear(x, w, b))
t_un = timeit(unfused)
t_hm = timeit(unfused_hm)
t_f = timeit(lambda: ops.qkv_self_attention(x, w, b, rows, mask))
print(f"rows={rows}: QKV Linear alone {t_lin:6.1f} us ({fl / t_lin / 1e6:5.0f} TF/s) | Linear + attention {t_un:6.1f} us | head-major "
      f"{t_hm:6.1f} us | fused {t_f:6.1f} us ({fl / t_f / 1e6:5.0f} TF/s of Linear flops)")
