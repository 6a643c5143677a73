import sys, os
sys.path.insert(0, "/root/repo")
import torch
from mraudio_b200 import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(M,N,K,bn,noflush=False):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev)*0.02).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    _lib.lib.mra_gemm_tile_override(bn)
    ts=[]
    for it in range(7):
        if not noflush: flush.zero_()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x,w,None,out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t=sorted(ts[1:])[3]
    return t*1e3, 2*M*N*K/t/1e9
for (M,N,K) in [(8192,768,768),(8192,768,1536),(8192,768,3072),(8192,768,6144),(8192,1536,3072),(8192,3072,3072),(16384,768,3072),(32768,768,3072),(8192,3072,768),(8192,768,3072)]:
    line=f"M={M} N={N} K={K}:"
    for bn in (128,192,256):
        t,tf=run(M,N,K,bn); line+=f" bn{bn} {t:6.1f}us {tf:5.0f}TF |"
    t,tf=run(M,N,K,256,noflush=True); line+=f" bn256-noflush {t:6.1f}us {tf:5.0f}TF"
    print(line, flush=True)
