"""Thin Python wrappers over the building-block entry points of the C-ABI (device tensors in, device tensors out).

These are what the parity tests call; the full forward goes through ``mraudio_b200.qformer``.
Every wrapper requires CUDA tensors and raises ``MraError`` otherwise -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, current_stream, lib, ptr


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MraError("mraudio_b200 ops need CUDA tensors (no CPU fallback)")


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, gelu: bool = False, out_fp32: bool = False,
           impl: int = _lib.GEMM_IMPL_TCGEN05, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = [gelu](x @ weight.T + bias) [+ residual];  x [M, K] bf16, weight [N, K] bf16, bias fp32 [N],
    residual fp32 [M, N];  y bf16 (or fp32 when out_fp32)."""
    _need_cuda(x, weight, bias, residual)
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
    assert x.dim() == 2 and weight.dim() == 2 and x.shape[1] == weight.shape[1]
    assert x.stride(1) == 1 and weight.stride(1) == 1
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.shape == (M, N) and residual.stride(1) == 1
    check(lib.mra_gemm_bf16(ptr(x), x.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(residual),
                            residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), M, N, K,
                            int(gelu), int(out_fp32), impl, current_stream()))
    return out


def wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dW[n_out, k_in] (+)= dy[n, n_out]^T @ x[n, k_in]; dy, x bf16 row-major; dW fp32."""
    _need_cuda(dy, x, out)
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dy.shape[0] == x.shape[0]
    assert dy.stride(1) == 1 and x.stride(1) == 1
    n, n_out = dy.shape
    k_in = x.shape[1]
    if out is None:
        assert not accumulate
        out = torch.empty(n_out, k_in, device=dy.device, dtype=torch.float32)
    check(lib.mra_wgrad_bf16(ptr(dy), dy.stride(0), ptr(x), x.stride(0), ptr(out), out.stride(0), n, n_out, k_in,
                             int(accumulate), current_stream()))
    return out


def dgrad(dy: torch.Tensor, weight: torch.Tensor, residual: Optional[torch.Tensor] = None, out_fp32: bool = False) -> torch.Tensor:
    """dX[n, k_in] = dy[n, n_out] @ weight[n_out, k_in] (+ residual): the Linear's weight is read as it lies."""
    _need_cuda(dy, weight, residual)
    assert dy.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and dy.shape[1] == weight.shape[0]
    assert dy.stride(1) == 1 and weight.stride(1) == 1
    n, n_out = dy.shape
    k_in = weight.shape[1]
    out = torch.empty(n, k_in, device=dy.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    check(lib.mra_dgrad_bf16(ptr(dy), dy.stride(0), ptr(weight), weight.stride(0), ptr(residual),
                             residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), n, n_out, k_in,
                             int(out_fp32), current_stream()))
    return out


def linear_residual_layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor,
                              gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    """(fp32, bf16) copies of LayerNorm(x @ weight.T + bias + residual) * gamma + beta; weight [768, K] bf16."""
    _need_cuda(x, weight, bias, residual, gamma, beta)
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and residual.dtype == torch.float32
    M, K = x.shape
    N = weight.shape[0]
    assert residual.shape == (M, N) and x.stride(1) == 1 and weight.stride(1) == 1 and residual.stride(1) == 1
    y32 = torch.empty(M, N, device=x.device, dtype=torch.float32)
    y16 = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_gemm_ln_bf16(ptr(x), x.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(residual), residual.stride(0),
                               ptr(gamma), ptr(beta), ptr(y32), N, ptr(y16), N, M, N, K, eps, current_stream()))
    return y32, y16


def split_residual(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 pair of the split residual stream: hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def linear_residual_layernorm_split(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], res_hi: torch.Tensor,
                                    res_lo: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                                    out: Optional[tuple] = None):
    """(hi, lo) bf16 pair of LayerNorm(x @ weight.T + bias + (res_hi + res_lo)) * gamma + beta; weight [768, K] bf16.
    `out` may alias the residual pair (the kernel reads every element before it writes it)."""
    _need_cuda(x, weight, bias, res_hi, res_lo, gamma, beta)
    assert x.dtype == weight.dtype == res_hi.dtype == res_lo.dtype == torch.bfloat16
    M, K = x.shape
    N = weight.shape[0]
    assert res_hi.shape == (M, N) and res_lo.shape == (M, N) and res_hi.stride() == res_lo.stride() and res_hi.stride(1) == 1
    assert x.stride(1) == 1 and weight.stride(1) == 1
    y_hi, y_lo = out if out is not None else (torch.empty(M, N, device=x.device, dtype=torch.bfloat16) for _ in range(2))
    check(lib.mra_gemm_ln_split_bf16(ptr(x), x.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(res_hi), ptr(res_lo),
                                     res_hi.stride(0), ptr(gamma), ptr(beta), ptr(y_hi), ptr(y_lo), y_hi.stride(0), M, N, K, eps,
                                     current_stream()))
    return y_hi, y_lo


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rows: int, heads: int, Sq: int, Sk: int,
              nq_split: int, kv_dense: bool, add_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [rows*Sq, >=heads*64] (split layout), k/v likewise or dense [rows*Sk, ...]; returns o [rows*Sq, heads*64]."""
    _need_cuda(q, k, v, add_mask)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    o = torch.empty(rows * Sq, heads * 64, device=q.device, dtype=torch.bfloat16)
    if add_mask is not None:
        assert add_mask.dtype == torch.float32 and add_mask.shape == (rows, Sk) and add_mask.is_contiguous()
    check(lib.mra_attention(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), o.stride(0),
                            ptr(add_mask), rows, heads, Sq, Sk, nq_split, int(kv_dense), current_stream()))
    return o


def linear_head_major(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x @ weight.T + bias written head-major: bf16 [N // 64, M, 64] (slot n // 64 holds columns 64*(n//64) .. of every row)."""
    _need_cuda(x, weight, bias)
    assert x.dtype == weight.dtype == torch.bfloat16 and x.stride(1) == 1 and weight.stride(1) == 1
    M, K = x.shape
    N = weight.shape[0]
    assert N % 64 == 0 and weight.shape[1] == K
    out = torch.empty(N // 64, M, 64, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_gemm_head_major_bf16(ptr(x), x.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(out), M, N, K,
                                       current_stream()))
    return out


def attention_head_major(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rows: int, Sq: int, Sk: int, nq_split: int,
                         kv_dense: bool, add_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q / k / v: head-major [heads, tokens, 64] (views of ``linear_head_major`` output); returns o [rows*Sq, heads*64]."""
    _need_cuda(q, k, v, add_mask)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    heads = q.shape[0]
    for t in (q, k, v):
        assert t.dim() == 3 and t.shape[0] == heads and t.shape[2] == 64 and t.stride(2) == 1
    o = torch.empty(rows * Sq, heads * 64, device=q.device, dtype=torch.bfloat16)
    if add_mask is not None:
        assert add_mask.dtype == torch.float32 and add_mask.shape == (rows, Sk) and add_mask.is_contiguous()
    check(lib.mra_attention_strided(ptr(q), q.stride(1), q.stride(0), ptr(k), k.stride(1), k.stride(0), ptr(v), v.stride(1),
                                    v.stride(0), ptr(o), o.stride(0), ptr(add_mask), rows, heads, Sq, Sk, nq_split,
                                    int(kv_dense), current_stream()))
    return o


def qkv_self_attention(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], rows: int,
                       add_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused QKV Linear + self-attention core: x [rows*64, K] in the split layout (32 query tokens of every row, then the 32
    text tokens of every row), weight [3*H, K]; returns ctx [rows*64, H] in the same row order."""
    _need_cuda(x, weight, bias, add_mask)
    assert x.dtype == weight.dtype == torch.bfloat16 and x.stride(1) == 1 and weight.stride(1) == 1
    M, K = x.shape
    H = weight.shape[0] // 3
    assert M == rows * 64 and weight.shape == (3 * H, K) and H % 64 == 0
    if add_mask is not None:
        assert add_mask.dtype == torch.float32 and add_mask.shape == (rows, 64) and add_mask.is_contiguous()
    ctx = torch.empty(M, H, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_qkv_attention_bf16(ptr(x), x.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(add_mask), ptr(ctx),
                                     ctx.stride(0), rows, H // 64, K, current_stream()))
    return ctx


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    """fp32 [rows, n] -> (fp32, bf16) LayerNorm outputs."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    y32 = torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_layernorm(ptr(x), ptr(gamma), ptr(beta), ptr(y32), ptr(y16), x.shape[0], x.shape[1], eps,
                            current_stream()))
    return y32, y16


_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def modality_layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
                       frame_major: bool = False) -> torch.Tensor:
    """x: [bs, F, Nk, W] (or [F, bs, Nk, W] when frame_major) -> bf16 [bs*F, Nk, W] batch-major rows."""
    _need_cuda(x, gamma, beta)
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in _DTYPE_CODE
    if frame_major:
        Fr, bs, Nk, W = x.shape
    else:
        bs, Fr, Nk, W = x.shape
    out = torch.empty(bs * Fr, Nk, W, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_modality_layernorm(ptr(x), _DTYPE_CODE[x.dtype], ptr(gamma), ptr(beta), ptr(out), bs, Fr, Nk, W,
                                     int(frame_major), eps, current_stream()))
    return out


def add_frame_position(x: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """x [bs, F, n, W] (fp32 / bf16 / fp16) + pos fp32 [>=F, W] broadcast over the n tokens -> bf16 [bs, F*n, W]."""
    _need_cuda(x, pos)
    assert x.dim() == 4 and x.is_contiguous() and x.dtype in _DTYPE_CODE
    bs, Fr, n, W = x.shape
    assert pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape[0] >= Fr and pos.shape[1] == W
    out = torch.empty(bs, Fr * n, W, device=x.device, dtype=torch.bfloat16)
    check(lib.mra_add_frame_position(ptr(x), _DTYPE_CODE[x.dtype], ptr(pos), ptr(out), bs, Fr, n, W, current_stream()))
    return out


def mr_score(pred: torch.Tensor, n_pred: torch.Tensor, gt: torch.Tensor, n_gt: torch.Tensor, thds: torch.Tensor):
    """pred f64 [Q, Pmax, 2], n_pred i32 [Q], gt f64 [Q, Gmax, 2], n_gt i32 [Q], thds f64 [10]
    -> (ap f64 [Q, 10], iou f64 [Q], invalid u8 [Q])."""
    _need_cuda(pred, n_pred, gt, n_gt, thds)
    Q, Pmax, _ = pred.shape
    Gmax = gt.shape[1]
    assert pred.dtype == gt.dtype == thds.dtype == torch.float64 and n_pred.dtype == n_gt.dtype == torch.int32
    assert pred.is_contiguous() and gt.is_contiguous() and thds.numel() == _lib.MRA_NUM_IOU_THDS
    ap = torch.empty(Q, _lib.MRA_NUM_IOU_THDS, device=pred.device, dtype=torch.float64)
    iou = torch.empty(Q, device=pred.device, dtype=torch.float64)
    inv = torch.empty(Q, device=pred.device, dtype=torch.uint8)
    check(lib.mra_mr_score(ptr(pred), ptr(n_pred), ptr(gt), ptr(n_gt), ptr(thds), Q, Pmax, Gmax, ptr(ap), ptr(iou),
                           ptr(inv), current_stream()))
    return ap, iou, inv
