"""Micro-benchmark of the tcgen05 GEMM on the shapes of the Q-Former step (tuning aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mraudio_b200 import ops, _lib

dev = torch.device("cuda:0")
if os.environ.get("GB_SHORT"):
    SHORT = {"kv_audio", "qkv", "f1x2", "f2x2", "proj", "f1_nogelu"}
else:
    SHORT = None
SHAPES = [  # name, M, N, K, gelu, f32+res
    ("kv_video", 65792, 9216, 1408, 0, 0), ("kv_audio", 65536, 9216, 768, 0, 0),
    ("qkv", 16384, 2304, 768, 0, 0), ("ao", 16384, 768, 768, 0, 1), ("cq", 8192, 768, 768, 0, 0),
    ("co", 8192, 768, 768, 0, 1), ("f1", 8192, 3072, 768, 1, 0), ("f2", 8192, 768, 3072, 0, 1),
    ("f1_nogelu", 8192, 3072, 768, 0, 0), ("f2_nores", 8192, 768, 3072, 0, 0), ("ao_nores", 16384, 768, 768, 0, 0),
    ("f1x2", 16384, 3072, 768, 1, 0), ("f2x2", 16384, 768, 3072, 0, 1), ("proj", 8192, 4096, 768, 0, 0),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, gelu, res in SHAPES:
    if SHORT is not None and name not in SHORT:
        continue
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if res else torch.bfloat16)
    line = f"[{os.environ.get('MRA_LIB', 'base')}] {name:9s} M={M:6d} N={N:5d} K={K:5d}"
    for bn in ((256, 0) if SHORT is not None else (-256, 128, 192, 256, 0)):
        _lib.lib.mra_gemm_cluster_override(1 if bn < 0 else int(os.environ.get('GB_CLUSTER', '3')))   # bn < 0: unpaired kernel at |bn| for comparison
        _lib.lib.mra_gemm_tile_override(abs(bn))
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.linear(x, w, b, residual=r, gelu=bool(gelu), out_fp32=bool(res), out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        line += f" | {'solo' if bn < 0 else 'bn'}{abs(bn):3d} {t*1e3:7.1f}us {2*M*N*K/t/1e9:6.0f}TF"
    _lib.lib.mra_gemm_tile_override(0)
    _lib.lib.mra_gemm_cluster_override(3)
    # cuBLAS (torch.matmul) for orientation only
    ts = []
    for it in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = x @ w.t(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    line += f" | cublas(no epi) {min(ts[1:])*1e3:7.1f}us"
    print(line, flush=True)

if SHORT is not None:
    sys.exit(0)
print("--- fused Linear + residual + LayerNorm (2-CTA cluster), N = 768")
for name, M, K in (("ao", 16384, 768), ("ao_x2", 32768, 768), ("co", 8192, 768), ("f2", 8192, 3072), ("f2_x4", 32768, 3072)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(768, K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(768, device=dev); r = torch.randn(M, 768, device=dev)
    g = torch.ones(768, device=dev); be = torch.zeros(768, device=dev)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear_residual_layernorm(x, w, b, r, g, be, 1e-12); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[2]
    print(f"{name:6s} M={M:6d} K={K:5d} fused {t*1e3:7.1f}us {2*M*768*K/t/1e9:6.0f}TF")
