"""A/B timing of the fused Linear + residual + LayerNorm kernel on the shapes of a config-2 step: fp32 residual stream vs the
split (bf16 hi + lo) stream.  L2 flushed before every launch; median of 7 after one warm-up.  Env: MRA_LIB=instr with
MRA_LN_DEBUG / MRA_LN_STAGGER selects the instrumented experiments (tools only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MRA_"))


def med(fn):
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[1:])
    return ts[len(ts) // 2] * 1e3


for name, M, K in (("ao_x2", 32768, 768), ("co_x2", 16384, 768), ("f2_x4", 32768, 3072)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(768, K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev)
    g = torch.ones(768, device=dev); be = torch.zeros(768, device=dev)
    rh, rl = ops.split_residual(r)
    oh, ol = torch.empty_like(rh), torch.empty_like(rl)
    t32 = med(lambda: ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12))
    tsp = med(lambda: ops.linear_residual_layernorm_split(x, w, b, rh, rl, g, be, 1e-12, out=(oh, ol)))
    # the same with the operand row pitch padded by 64 elements (128 B): does the power-of-two-ish pitch of K = 3072 camp on
    # L2 slices / DRAM channels?
    xp = torch.empty(M, K + 64, device=dev, dtype=torch.bfloat16)[:, :K]; xp.copy_(x)
    wp = torch.empty(768, K + 64, device=dev, dtype=torch.bfloat16)[:, :K]; wp.copy_(w)
    tpad = med(lambda: ops.linear_residual_layernorm_split(xp, wp, b, rh, rl, g, be, 1e-12, out=(oh, ol)))
    tplain = med(lambda: ops.linear(x, w, b))
    tplain_pad = med(lambda: ops.linear(xp, wp, b))
    fl = 2 * M * 768 * K
    b32 = M * (K * 2 + 768 * (4 + 4 + 2)); bsp = M * (K * 2 + 768 * (4 + 4))
    print(f"[{tag}] {name:6s} M={M:6d} K={K:5d} | fp32 stream {t32:7.1f} us {fl/t32/1e6:6.0f} TF/s {b32/t32/1e3:6.0f} GB/s"
          f" | split stream {tsp:7.1f} us {fl/tsp/1e6:6.0f} TF/s {bsp/tsp/1e3:6.0f} GB/s | split, padded pitch {tpad:7.1f} us"
          f" | plain GEMM bf16 out {tplain:7.1f} us, padded pitch {tplain_pad:7.1f} us", flush=True)
