"""GPU parity of the building-block kernels, called through the C-ABI (mraudio_b200.ops -> libmraudio_b200.so)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


GEMM_SHAPES = [
    (128, 128, 64), (128, 256, 64), (256, 768, 768), (1000, 2304, 768), (257 * 3, 1536, 1408), (64, 4096, 768),
    (8192, 3072, 768), (640, 768, 3072), (33, 8, 72), (129, 264, 200), (1, 768, 768),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
def test_gemm_matches_fp32_reference(M, N, K, impl):
    from mraudio_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    ref = x.float() @ w.float().t() + b
    y = ops.linear(x, w, b, out_fp32=True, impl=impl)
    torch.cuda.synchronize()
    assert _rel(y, ref) < 2e-5 * math.sqrt(K), "fp32-out GEMM must match an fp32 matmul of the same bf16 operands"
    y16 = ops.linear(x, w, b, impl=impl)
    assert _rel(y16, ref) < 6e-3


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
def test_gemm_epilogues(impl):
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, N, K = 300, 768, 3072
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.02).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    res = torch.randn(M, N, generator=g).to(_dev())
    lin = x.float() @ w.float().t() + b
    assert _rel(ops.linear(x, w, b, residual=res, out_fp32=True, impl=impl), lin + res) < 2e-3
    assert _rel(ops.linear(x, w, b, gelu=True, out_fp32=True, impl=impl), torch.nn.functional.gelu(lin)) < 2e-3
    assert _rel(ops.linear(x, w, None, out_fp32=True, impl=impl), lin - b) < 2e-3
    # every epilogue combination incl. bf16 output + residual and GELU + residual
    for gelu in (False, True):
        for f32 in (False, True):
            ref = torch.nn.functional.gelu(lin) if gelu else lin
            got = ops.linear(x, w, b, residual=res, gelu=gelu, out_fp32=f32, impl=impl)
            assert _rel(got, ref + res) < (2e-3 if f32 else 6e-3), (gelu, f32)
    # the fast erf of the tensor-core epilogue is accurate far below bf16 resolution
    y = ops.linear(x, w, b, gelu=True, out_fp32=True, impl=impl)
    assert (y - torch.nn.functional.gelu(ops.linear(x, w, b, out_fp32=True, impl=impl))).abs().max().item() < 2e-6
    # strided operands (a column slice of a wider matrix) and a strided output
    big = torch.randn(M, 2 * K, generator=g).to(_dev(), torch.bfloat16)
    xs = big[:, K:]
    out = torch.zeros(M, 2 * N, device=_dev(), dtype=torch.bfloat16)
    ops.linear(xs, w, b, impl=impl, out=out[:, N:])
    assert _rel(out[:, N:], xs.float() @ w.float().t() + b) < 6e-3
    assert out[:, :N].abs().max().item() == 0.0


@pytest.mark.parametrize("bn", [128, 192, 256])
@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (1000, 2304, 768), (129, 264, 200), (4096, 3072, 768)])
def test_gemm_every_tile_width(bn, M, N, K):
    from mraudio_b200 import ops, _lib
    g = torch.Generator().manual_seed(bn + M)
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    res = torch.randn(M, N, generator=g).to(_dev())
    ref = x.float() @ w.float().t() + b
    try:
        _lib.check(_lib.lib.mra_gemm_tile_override(bn))
        assert _rel(ops.linear(x, w, b, residual=res, out_fp32=True), ref + res) < 1e-4
        assert _rel(ops.linear(x, w, b, gelu=True), torch.nn.functional.gelu(ref)) < 6e-3
        assert _rel(ops.linear(x, w, b), ref) < 6e-3
    finally:
        _lib.lib.mra_gemm_tile_override(0)


@pytest.mark.parametrize("M,K", [(128, 768), (256, 768), (1000, 768), (8192, 768), (4096, 3072), (130, 64), (1, 768), (33000, 768)])
def test_fused_linear_residual_layernorm(M, K):
    """2-CTA-cluster GEMM + bias + residual + LayerNorm (N = 768) against fp32 PyTorch on the same bf16 operands."""
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    N = 768
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.03).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    res = (torch.randn(M, N, generator=g) * 2 + 0.3).to(_dev())
    gam = (1 + 0.2 * torch.randn(N, generator=g)).to(_dev())
    bet = (0.3 * torch.randn(N, generator=g)).to(_dev())
    pre = x.float() @ w.float().t() + b + res
    ref = torch.nn.functional.layer_norm(pre, (N,), gam, bet, 1e-12)
    y32, y16 = ops.linear_residual_layernorm(x, w, b, res, gam, bet, 1e-12)
    torch.cuda.synchronize()
    assert (y32 - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    assert torch.equal(y16, y32.to(torch.bfloat16))
    # no bias; deterministic
    y32b, _ = ops.linear_residual_layernorm(x, w, None, res, gam, bet, 1e-12)
    ref_b = torch.nn.functional.layer_norm(pre - b, (N,), gam, bet, 1e-12)
    assert (y32b - ref_b).abs().max().item() < 2e-4 * max(1.0, ref_b.abs().max().item())
    y32c, _ = ops.linear_residual_layernorm(x, w, b, res, gam, bet, 1e-12)
    assert torch.equal(y32, y32c)


@pytest.mark.parametrize("M,K", [(128, 768), (256, 768), (1000, 768), (8192, 768), (4096, 3072), (130, 64), (1, 768), (33000, 768)])
def test_fused_linear_residual_layernorm_split(M, K):
    """The same kernel with the residual stream as a bf16 (hi, lo) pair -- the form the inference forward uses: hi + lo must
    match the fp32 LayerNorm of (x W^T + b + (res_hi + res_lo)) to 16 mantissa bits, hi must be exactly bf16(hi + lo)'s
    leading part (it is the next GEMM's operand), in-place operation on the residual pair must give the same bits, and
    rows outside [0, M) must stay untouched."""
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(3 * M + K)
    N = 768
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.03).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    res = (torch.randn(M, N, generator=g) * 2 + 0.3).to(_dev())
    gam = (1 + 0.2 * torch.randn(N, generator=g)).to(_dev())
    bet = (0.3 * torch.randn(N, generator=g)).to(_dev())
    r_hi, r_lo = ops.split_residual(res)
    res_q = r_hi.float() + r_lo.float()
    assert (res_q - res).abs().max().item() <= 2.0 ** -16 * res.abs().max().item()
    pre = x.float() @ w.float().t() + b + res_q
    ref = torch.nn.functional.layer_norm(pre, (N,), gam, bet, 1e-12)
    pad = 32
    bh = torch.full((M + 2 * pad, N), 7.0, device=_dev(), dtype=torch.bfloat16)
    bl = torch.full((M + 2 * pad, N), 7.0, device=_dev(), dtype=torch.bfloat16)
    y_hi, y_lo = ops.linear_residual_layernorm_split(x, w, b, r_hi, r_lo, gam, bet, 1e-12, out=(bh[pad:pad + M], bl[pad:pad + M]))
    torch.cuda.synchronize()
    for t in (bh, bl):
        assert (t[:pad] == 7.0).all() and (t[pad + M:] == 7.0).all()
    y = y_hi.float() + y_lo.float()
    scale = max(1.0, ref.abs().max().item())
    assert (y - ref).abs().max().item() < 2e-4 * scale
    # hi is the bf16 rounding of the fp32 result the kernel held (|lo| <= half an ulp of hi)
    assert (y_lo.float().abs() <= y_hi.float().abs() * 2.0 ** -8 + 1e-30).all()
    # in place on the residual pair (cross-attention output block of the forward): identical bits
    h2, l2 = r_hi.clone(), r_lo.clone()
    ops.linear_residual_layernorm_split(x, w, b, h2, l2, gam, bet, 1e-12, out=(h2, l2))
    torch.cuda.synchronize()
    assert torch.equal(h2, y_hi) and torch.equal(l2, y_lo)
    # against the fp32-stream kernel on the same (rounded) residual: same arithmetic up to the output rounding
    # (the two forms sum the row statistics in a different order, so the last bit of mean / rstd may differ)
    y32, y16 = ops.linear_residual_layernorm(x, w, b, res_q, gam, bet, 1e-12)
    assert (y32 - y).abs().max().item() <= 2.0 ** -14 * scale
    assert (y16.float() - y_hi.float()).abs().max().item() <= 2.0 ** -7 * scale   # at most one bf16 ulp, where a rounding flips
    assert (y16 != y_hi).float().mean().item() < 1e-2


@pytest.mark.parametrize("cm", [1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(512, 768, 768), (771, 2304, 768), (1280, 3072, 768), (8192, 768, 3072), (640, 9216, 1408)])
def test_gemm_cluster_multicast_variant(cm, M, N, K):
    """Pairs of CTAs along M: cm = 2 shares the W slab by TMA multicast, cm = 3 runs one tcgen05.mma.cta_group::2 (M = 256, half
    of the W slab per CTA); both == unpaired kernel (cm = 1) == fp32 reference, including an odd number of row blocks."""
    from mraudio_b200 import ops, _lib
    g = torch.Generator().manual_seed(cm * 1000 + M)
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    res = torch.randn(M, N, generator=g).to(_dev())
    ref = x.float() @ w.float().t() + b
    try:
        _lib.check(_lib.lib.mra_gemm_cluster_override(cm))
        for bn in (0, 128, 192, 256):
            _lib.check(_lib.lib.mra_gemm_tile_override(bn))
            assert _rel(ops.linear(x, w, b, residual=res, out_fp32=True), ref + res) < 1e-4, bn
            assert _rel(ops.linear(x, w, b, gelu=True), torch.nn.functional.gelu(ref)) < 6e-3, bn
    finally:
        _lib.lib.mra_gemm_cluster_override(3)
        _lib.lib.mra_gemm_tile_override(0)


@pytest.mark.parametrize("M,N,K", [(771, 2304, 768), (1281, 768, 3072), (640, 776, 768)])
def test_paired_kernels_write_nothing_outside_their_output(M, N, K):
    """Guard rows around the outputs of the 2-CTA-MMA GEMM and of the fused GEMM+LayerNorm (6-CTA clusters) with an odd
    number of 128-row blocks and a ragged last tile: the TMA stores must clip at M / N exactly."""
    from mraudio_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    pad = 64
    buf = torch.full((M + 2 * pad, N), 7.0, device=_dev(), dtype=torch.bfloat16)
    out = buf[pad:pad + M]
    _lib.check(_lib.lib.mra_gemm_bf16(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(b), None, 0, _lib.ptr(out), N, M, N, K, 0, 0, 0,
                                      _lib.current_stream()))
    torch.cuda.synchronize()
    assert (buf[:pad] == 7.0).all() and (buf[pad + M:] == 7.0).all()
    assert _rel(out, x.float() @ w.float().t() + b) < 6e-3
    if N == 768:
        res = torch.randn(M, N, generator=g).to(_dev())
        gam, bet = torch.ones(N, device=_dev()), torch.zeros(N, device=_dev())
        b32 = torch.full((M + 2 * pad, N), 7.0, device=_dev(), dtype=torch.float32)
        b16 = torch.full((M + 2 * pad, N), 7.0, device=_dev(), dtype=torch.bfloat16)
        _lib.check(_lib.lib.mra_gemm_ln_bf16(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(res), N, _lib.ptr(gam), _lib.ptr(bet),
                                             _lib.ptr(b32[pad:]), N, _lib.ptr(b16[pad:]), N, M, N, K, 1e-12, _lib.current_stream()))
        torch.cuda.synchronize()
        for t in (b32, b16):
            assert (t[:pad] == 7.0).all() and (t[pad + M:] == 7.0).all()
        ref = torch.nn.functional.layer_norm(x.float() @ w.float().t() + b + res, (N,), gam, bet, 1e-12)
        assert _rel(b32[pad:pad + M], ref) < 2e-3


@pytest.mark.parametrize("n,n_out,k_in", [(64, 128, 64), (4096, 768, 768), (2048, 3072, 768), (2048, 768, 3072), (1000, 2304, 768),
                                          (16448, 1536, 1408), (77, 72, 200), (130, 4096, 768)])
def test_wgrad_mn_major_operands(n, n_out, k_in):
    """dW = dY^T X straight from the row-major dY / X (MN-major UMMA descriptors), incl. in-place accumulation."""
    from mraudio_b200 import ops, _lib
    g = torch.Generator().manual_seed(n + n_out)
    dy = torch.randn(n, n_out, generator=g).to(_dev(), torch.bfloat16)
    x = torch.randn(n, k_in, generator=g).to(_dev(), torch.bfloat16)
    ref = dy.float().t() @ x.float()
    try:
        for bn in (0, 128, 192, 256):
            _lib.check(_lib.lib.mra_gemm_tile_override(bn))
            dw = ops.wgrad(dy, x)
            assert _rel(dw, ref) < 1e-4, bn
            acc = torch.full_like(dw, 0.5)
            ops.wgrad(dy, x, out=acc, accumulate=True)
            assert _rel(acc, ref + 0.5) < 1e-4, bn
    finally:
        _lib.lib.mra_gemm_tile_override(0)
    # strided views (a column block of a wider matrix, as the backward passes them)
    big = torch.randn(n, n_out + 64, generator=g).to(_dev(), torch.bfloat16)
    dw = ops.wgrad(big[:, 64:], x)
    assert _rel(dw, big[:, 64:].float().t() @ x.float()) < 1e-4


@pytest.mark.parametrize("n,n_out,k_in", [(64, 64, 64), (4096, 768, 768), (2048, 3072, 768), (2048, 768, 3072), (1000, 2304, 768),
                                          (130, 4096, 768), (77, 72, 200)])
def test_dgrad_reads_forward_weight_in_place(n, n_out, k_in):
    """dX = dY W straight from the forward's [n_out, k_in] weight (MN-major descriptor for the second operand)."""
    from mraudio_b200 import ops, _lib
    g = torch.Generator().manual_seed(n + k_in)
    dy = torch.randn(n, n_out, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(n_out, k_in, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    res = torch.randn(n, k_in, generator=g).to(_dev())
    ref = dy.float() @ w.float()
    try:
        for bn in (0, 128, 192, 256):
            _lib.check(_lib.lib.mra_gemm_tile_override(bn))
            assert _rel(ops.dgrad(dy, w), ref) < 6e-3, bn
            assert _rel(ops.dgrad(dy, w, out_fp32=True), ref) < 1e-4, bn
            assert _rel(ops.dgrad(dy, w, residual=res, out_fp32=True), ref + res) < 1e-4, bn
    finally:
        _lib.lib.mra_gemm_tile_override(0)


def test_gemm_tcgen05_equals_simt_bitwise_ordering_free():
    """Same bf16 operands, fp32 accumulation: the two implementations agree to fp32 rounding noise."""
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(777, 1408, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(1536, 1408, generator=g) * 0.02).to(_dev(), torch.bfloat16)
    a = ops.linear(x, w, out_fp32=True, impl=0)
    b = ops.linear(x, w, out_fp32=True, impl=1)
    assert _rel(a, b) < 1e-5


@pytest.fixture(params=["tma", "generic"])
def attn_impl(request):
    from mraudio_b200 import _lib
    _lib.lib.mra_attention_impl_override(1 if request.param == "generic" else 0)
    yield request.param
    _lib.lib.mra_attention_impl_override(0)


def _ref_attention(q, k, v, mask, heads):
    R, Sq, H = q.shape
    Sk = k.shape[1]
    qh = q.float().view(R, Sq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(R, Sk, heads, 64).permute(0, 2, 1, 3)
    vh = v.float().view(R, Sk, heads, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + mask[:, None, None, :]
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(R, Sq, H)


@pytest.mark.parametrize("rows,Sq,Sk,heads", [(3, 32, 257, 12), (2, 32, 256, 12), (5, 32, 8, 12), (2, 32, 1024, 12),
                                               (1, 8, 8, 2), (4, 17, 100, 3), (2, 64, 300, 4), (2, 200, 70, 2)])
def test_cross_attention(rows, Sq, Sk, heads, attn_impl):
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(rows * 100 + Sk)
    H = heads * 64
    q = torch.randn(rows, Sq, H, generator=g).to(_dev(), torch.bfloat16)
    kv = torch.randn(rows, Sk, 2 * H, generator=g).to(_dev(), torch.bfloat16)
    mask = torch.zeros(rows, Sk)
    mask[0, Sk // 2:] = -10000.0
    mask = mask.to(_dev())
    kvf = kv.view(rows * Sk, 2 * H)
    for m in (None, mask):
        o = ops.attention(q.view(rows * Sq, H), kvf[:, :H], kvf[:, H:], rows, heads, Sq, Sk, Sq, True, m)
        ref = _ref_attention(q, kv[..., :H], kv[..., H:], m, heads)
        assert _rel(o.view(rows, Sq, H), ref) < 1.5e-2


@pytest.mark.parametrize("rows,Nq,T", [(3, 32, 32), (2, 32, 0), (4, 32, 13), (2, 32, 128), (2, 8, 0), (3, 64, 40), (2, 8, 5)])
def test_self_attention_split_layout(rows, Nq, T, attn_impl):
    from mraudio_b200 import ops
    heads, H = 12, 768
    S = Nq + T
    g = torch.Generator().manual_seed(rows + T)
    qkv = torch.randn(rows, S, 3 * H, generator=g).to(_dev(), torch.bfloat16)
    mask = torch.zeros(rows, S)
    if T > 3:
        mask[1, Nq + T // 2:] = -10000.0
    mask = mask.to(_dev())
    # split layout: all query tokens first, then all text tokens
    split = torch.cat([qkv[:, :Nq].reshape(rows * Nq, 3 * H), qkv[:, Nq:].reshape(rows * T, 3 * H)], 0).contiguous()
    o = ops.attention(split[:, :H], split[:, H:2 * H], split[:, 2 * H:], rows, heads, S, S, Nq, False, mask)
    ref = _ref_attention(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], mask, heads)
    got = torch.cat([o[:rows * Nq].view(rows, Nq, H), o[rows * Nq:].view(rows, T, H)], 1)
    assert _rel(got, ref) < 1.5e-2


@pytest.mark.parametrize("M,N,K", [(32 * 257, 1536, 1408), (777, 2304, 768), (100, 64, 64), (4096, 9216, 768)])
def test_gemm_head_major_output(M, N, K):
    """The head-major epilogue ([N / 64][M][64], what the inference forward hands to the attention kernel) holds exactly
    the values of the row-major bf16 GEMM: same accumulation, only the TMA store coordinates differ."""
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(_dev(), torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.03).to(_dev(), torch.bfloat16)
    b = torch.randn(N, generator=g).to(_dev())
    guard = torch.full((N // 64 + 2, M, 64), 7.0, device=_dev(), dtype=torch.bfloat16)   # slots 0 and -1 must stay untouched
    hm = ops.linear_head_major(x, w, b)
    ref = ops.linear(x, w, b)
    assert torch.equal(hm.permute(1, 0, 2).reshape(M, N), ref)
    from mraudio_b200 import _lib
    _lib.check(_lib.lib.mra_gemm_head_major_bf16(_lib.ptr(x), x.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(b),
                                                 guard[1:].data_ptr(), M, N, K, _lib.current_stream()))
    assert torch.equal(guard[1:-1], hm) and bool((guard[0] == 7.0).all()) and bool((guard[-1] == 7.0).all())


@pytest.mark.parametrize("rows,Sq,Sk,heads", [(3, 32, 257, 12), (2, 32, 256, 12), (4, 17, 100, 3), (2, 64, 300, 4), (5, 32, 8, 12)])
def test_cross_attention_head_major_equals_row_major(rows, Sq, Sk, heads, attn_impl):
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(rows * 10 + Sk)
    H = heads * 64
    q = torch.randn(rows * Sq, H, generator=g).to(_dev(), torch.bfloat16)
    kv = torch.randn(rows * Sk, 2 * H, generator=g).to(_dev(), torch.bfloat16)
    mask = torch.zeros(rows, Sk)
    mask[0, Sk // 2:] = -10000.0
    mask = mask.to(_dev())
    ref = ops.attention(q, kv[:, :H], kv[:, H:], rows, heads, Sq, Sk, Sq, True, mask)
    kv_hm = kv.view(rows * Sk, 2 * heads, 64).permute(1, 0, 2).contiguous()       # [2 * heads][tokens][64]
    q_hm = q.view(rows * Sq, heads, 64).permute(1, 0, 2).contiguous()
    o = ops.attention_head_major(q_hm, kv_hm[:heads], kv_hm[heads:], rows, Sq, Sk, Sq, True, mask)
    assert torch.equal(o, ref)
    # mixed: row-major queries (stride 64 between heads), head-major keys / values -- the cross-attention of the forward
    o2 = ops.attention_head_major(q.view(rows * Sq, heads, 64).permute(1, 0, 2), kv_hm[:heads], kv_hm[heads:], rows, Sq, Sk, Sq,
                                  True, mask)
    assert torch.equal(o2, ref)


@pytest.mark.parametrize("rows,Nq,T", [(3, 32, 32), (2, 32, 0), (4, 32, 13), (2, 32, 128), (3, 64, 40)])
def test_self_attention_head_major_equals_row_major(rows, Nq, T, attn_impl):
    from mraudio_b200 import ops
    heads, H = 12, 768
    S = Nq + T
    g = torch.Generator().manual_seed(rows + 3 * T)
    split = torch.randn(rows * S, 3 * H, generator=g).to(_dev(), torch.bfloat16)   # split layout: query rows, then text rows
    mask = torch.zeros(rows, S)
    if T > 3:
        mask[1, Nq + T // 2:] = -10000.0
    mask = mask.to(_dev())
    ref = ops.attention(split[:, :H], split[:, H:2 * H], split[:, 2 * H:], rows, heads, S, S, Nq, False, mask)
    hm = split.view(rows * S, 3 * heads, 64).permute(1, 0, 2).contiguous()         # [3 * heads][tokens][64]
    o = ops.attention_head_major(hm[:heads], hm[heads:2 * heads], hm[2 * heads:], rows, S, S, Nq, False, mask)
    assert torch.equal(o, ref)


@pytest.mark.parametrize("rows,heads,K,masked", [(1, 12, 768, False), (2, 12, 768, True), (3, 12, 768, True), (4, 12, 768, False),
                                                 (5, 4, 256, True), (37, 12, 768, True), (256, 12, 768, True)])
def test_fused_qkv_attention_equals_linear_then_attention(rows, heads, K, masked):
    """csrc/qkv_attn.cu: the QKV Linear and the self-attention core in one kernel (Q / K / V stay in shared memory) computes
    exactly what the two launches it replaces compute: same fp32 accumulation + bias + bf16 rounding of q / k / v, same
    flash step.  Covers clip blocks cut at the tail (rows not a multiple of 4 / 2), a second geometry, and the text mask."""
    from mraudio_b200 import ops
    H = heads * 64
    g = torch.Generator().manual_seed(rows * 7 + heads)
    x = torch.randn(rows * 64, K, generator=g).to(_dev(), torch.bfloat16)          # split layout: query rows, then text rows
    w = (torch.randn(3 * H, K, generator=g) * 0.05).to(_dev(), torch.bfloat16)
    b = (torch.randn(3 * H, generator=g) * 0.5).to(_dev())
    mask = None
    if masked:
        mask = torch.zeros(rows, 64)
        mask[rows // 2, 32 + 11:] = -10000.0
        mask[0, 60:] = -10000.0
        mask = mask.to(_dev())
    guard = torch.full((rows * 64 + 16, H), 3.0, device=_dev(), dtype=torch.bfloat16)
    qkv = ops.linear(x, w, b)
    ref = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], rows, heads, 64, 64, 32, False, mask)
    got = ops.qkv_self_attention(x, w, b, rows, mask)
    assert torch.equal(got, ref)
    # nothing written outside [rows * 64, H], also when the last clip block is cut
    from mraudio_b200 import _lib
    out = guard[8:8 + rows * 64]
    _lib.check(_lib.lib.mra_qkv_attention_bf16(_lib.ptr(x), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(mask), out.data_ptr(), H, rows,
                                               heads, K, _lib.current_stream()))
    assert torch.equal(out, ref) and bool((guard[:8] == 3.0).all()) and bool((guard[8 + rows * 64:] == 3.0).all())
    # against the fp32 definition
    xs = torch.cat([x[:rows * 32].view(rows, 32, K), x[rows * 32:].view(rows, 32, K)], 1).float()
    full = xs @ w.float().t() + b
    r32 = _ref_attention(full[..., :H], full[..., H:2 * H], full[..., 2 * H:], mask, heads)
    got3 = torch.cat([got[:rows * 32].view(rows, 32, H), got[rows * 32:].view(rows, 32, H)], 1)
    assert _rel(got3, r32) < 1.5e-2


def test_attention_unaligned_output_rows_take_the_narrow_store_path():
    """Output rows that are not 16-byte aligned cannot leave as 128-byte lines: same values through the 4-byte stores."""
    from mraudio_b200 import _lib, ops
    rows, Sq, Sk, heads, H = 2, 32, 257, 2, 128
    g = torch.Generator().manual_seed(5)
    q = torch.randn(rows * Sq, H, generator=g).to(_dev(), torch.bfloat16)
    kv = torch.randn(rows * Sk, 2 * H, generator=g).to(_dev(), torch.bfloat16)
    ref = ops.attention(q, kv[:, :H], kv[:, H:], rows, heads, Sq, Sk, Sq, True, None)
    buf = torch.zeros(rows * Sq, H + 2, device=_dev(), dtype=torch.bfloat16)
    o = buf[:, 2:]                                       # rows start 4 bytes off a 16-byte boundary, pitch 260 bytes
    _lib.check(_lib.lib.mra_attention(_lib.ptr(q), H, _lib.ptr(kv), 2 * H, kv[:, H:].data_ptr(), 2 * H, o.data_ptr(), H + 2, None,
                                      rows, heads, Sq, Sk, Sq, 1, _lib.current_stream()))
    assert torch.equal(o, ref) and bool((buf[:, :2] == 0).all())


@pytest.mark.parametrize("rows,n", [(1, 768), (1000, 768), (77, 1408), (5, 1024), (9, 8)])
def test_layernorm(rows, n):
    from mraudio_b200 import ops
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(rows, n, generator=g) * 3 + 1).to(_dev())
    gam = torch.randn(n, generator=g).to(_dev())
    bet = torch.randn(n, generator=g).to(_dev())
    y32, y16 = ops.layernorm(x, gam, bet, 1e-12)
    ref = torch.nn.functional.layer_norm(x, (n,), gam, bet, 1e-12)
    assert (y32 - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-6
    assert torch.equal(y16, y32.to(torch.bfloat16))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("frame_major", [False, True])
def test_modality_layernorm_and_frame_fold(dtype, frame_major):
    """models/xinstructblip.py:822-828 (fp32-upcast LN) + :280-285 (frame fold, batch-major reorder)."""
    from mraudio_b200 import ops
    bs, Fr, Nk, W = 3, 4, 17, 1408
    g = torch.Generator().manual_seed(2)
    x = torch.randn(bs, Fr, Nk, W, generator=g).to(dtype)
    gam = (1 + 0.1 * torch.randn(W, generator=g)).to(_dev())
    bet = (0.1 * torch.randn(W, generator=g)).to(_dev())
    ref = torch.nn.functional.layer_norm(x.float().to(_dev()), (W,), gam, bet, 1e-5).reshape(bs * Fr, Nk, W)
    xin = x.permute(1, 0, 2, 3).contiguous() if frame_major else x
    y = ops.modality_layernorm(xin.to(_dev()), gam, bet, 1e-5, frame_major=frame_major)
    assert y.shape == (bs * Fr, Nk, W) and y.dtype == torch.bfloat16
    assert (y.float() - ref).abs().max().item() < 4e-2  # bf16 output rounding of values up to ~5
    assert _rel(y, ref) < 5e-3
