"""Parity of the tcgen05 implicit-GEMM operator (sonic_conv_gemm) against fp32 PyTorch.

The fp32 torch reference here is the per-operator form of the oracle UNet's layers
(oracle/unet.py: nn.Conv2d / nn.Linear / GEGLU); inputs are rounded to bf16 first so the only
differences are accumulation order and the single bf16 rounding of the output.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel_err(got, ref):
    return ((got.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 320, 320), (77 * 2, 640, 768), (1000, 1280, 1280),
                                    (4096, 960, 320), (77, 512, 512), (29, 512, 512), (1, 768, 512)])
def test_linear(cuda, M, N, K):
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device=cuda, generator=g))
    w = _bf(torch.randn(N, K, device=cuda, generator=g) / K ** 0.5)
    b = torch.randn(N, device=cuda, generator=g)
    res = _bf(torch.randn(M, N, device=cuda, generator=g))
    out = k.conv_gemm(a, w, N, bias=b, residual=res)
    ref = a.float() @ w.float().t() + b + res.float()
    torch.cuda.synchronize()
    assert _rel_err(out, ref) < 1e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (3, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                             (3, 8, 8, 1280, 1280), (1, 8, 8, 64, 64)])
def test_conv3x3(cuda, B, H, W, Cin, Cout):
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(B * H + Cin)
    x = _bf(torch.randn(B, H, W, Cin, device=cuda, generator=g))
    w = _bf(torch.randn(Cout, Cin, 3, 3, device=cuda, generator=g) / (9 * Cin) ** 0.5)
    b = torch.randn(Cout, device=cuda, generator=g)
    rb = torch.randn(B, Cout, device=cuda, generator=g)
    out = k.conv_gemm(x, k.pack_conv3x3_weight(w), Cout, taps=9, n_img=B, H=H, W=W, bias=b, row_bias=rb)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1) + rb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    torch.cuda.synchronize()
    assert _rel_err(out, ref) < 1e-2


def test_conv3x3_concat_and_1x1(cuda):
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, W, C0, C1, Cout = 2, 32, 32, 640, 320, 640
    x0 = _bf(torch.randn(B, H, W, C0, device=cuda, generator=g))
    x1 = _bf(torch.randn(B, H, W, C1, device=cuda, generator=g))
    w = _bf(torch.randn(Cout, C0 + C1, 3, 3, device=cuda, generator=g) / (9 * (C0 + C1)) ** 0.5)
    out = k.conv_gemm(x0, k.pack_conv3x3_weight(w), Cout, taps=9, n_img=B, H=H, W=W, a1=x1)
    xc = torch.cat([x0, x1], dim=-1).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xc, w.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert _rel_err(out, ref) < 1e-2
    w1 = _bf(torch.randn(Cout, C0 + C1, device=cuda, generator=g) / (C0 + C1) ** 0.5)
    out1 = k.conv_gemm(x0, w1, Cout, taps=1, n_img=B, H=H, W=W, a1=x1)
    ref1 = torch.cat([x0, x1], dim=-1).float().reshape(-1, C0 + C1) @ w1.float().t()
    assert _rel_err(out1, ref1) < 1e-2


@pytest.mark.parametrize("M,C", [(4096, 320), (256, 1280)])
def test_geglu(cuda, M, C):
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(C)
    x = _bf(torch.randn(M, C, device=cuda, generator=g))
    w = _bf(torch.randn(8 * C, C, device=cuda, generator=g) / C ** 0.5)
    b = torch.randn(8 * C, device=cuda, generator=g)
    bn = k.gemm_block_n(8 * C, 1, 1, M, k.EPI_GEGLU)
    wp, bp = k.pack_geglu(w, b, bn)
    out = k.conv_gemm(x, wp, 8 * C, bias=bp, epilogue=k.EPI_GEGLU, block_n=bn)
    h = x.float() @ w.float().t() + b
    ref = h[:, :4 * C] * F.gelu(h[:, 4 * C:])
    assert out.shape == (M, 4 * C)
    assert _rel_err(out, ref) < 1e-2


def test_residual_in_place_and_ragged_rows(cuda):
    """out aliases residual (the transformer blocks update h in place) and M is not a multiple of the 128-row tile /
    32-row staging chunk: the TMA-store epilogue must clip, never write past row M."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 1000 + 13, 320, 640
    a = _bf(torch.randn(M, K, device=cuda, generator=g))
    w = _bf(torch.randn(N, K, device=cuda, generator=g) / K ** 0.5)
    b = torch.randn(N, device=cuda, generator=g)
    buf = _bf(torch.randn(M + 64, N, device=cuda, generator=g))       # guard rows behind the matrix
    h = buf[:M]
    guard = buf[M:].clone()
    ref = a.float() @ w.float().t() + b + h.float()
    k.conv_gemm(a, w, N, bias=b, residual=h, out=h)
    torch.cuda.synchronize()
    assert _rel_err(h, ref) < 1e-2
    assert torch.equal(buf[M:], guard), "epilogue wrote past the last row"


def test_narrow_and_strided_outputs(cuda):
    """N=16 (the padded conv_out) takes the direct-store epilogue; a column slice of a wider buffer (ld_out > N) and a
    non-multiple-of-32 N tail take the staged one."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(12)
    for M, N, K in [(4096, 16, 320), (512, 80, 128), (640, 352, 256)]:
        a = _bf(torch.randn(M, K, device=cuda, generator=g))
        w = _bf(torch.randn(N, K, device=cuda, generator=g) / K ** 0.5)
        wide = torch.zeros(M, N + 64, device=cuda, dtype=torch.bfloat16)
        out = wide[:, 32:32 + N]
        k.conv_gemm(a, w, N, out=out)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        assert _rel_err(out, ref) < 1e-2, (M, N, K)
        assert wide[:, :32].abs().max().item() == 0 and wide[:, 32 + N:].abs().max().item() == 0, (M, N, K)


@pytest.mark.parametrize("M,C,N,geglu", [(4096, 320, 960, False), (1000, 640, 640, False), (512, 1280, 3840, False),
                                          (4096, 320, 2560, True), (300, 1280, 10240, True)])
def test_layernorm_folded_into_gemm(cuda, M, C, N, geglu):
    """The LayerNorms of the transformer blocks launch no kernel: the PRODUCER GEMM (here a +residual projection
    of width C) leaves per-row (sum, sumsq) partials of its output, the CONSUMER (a plain / GEGLU projection)
    takes (-mean, std) through the tensor core as one extra K chunk (``sonic_ln_side``'s side tensor against the
    (s, b') columns of ``fold_layernorm``) and multiplies by rstd in its epilogue.  Reference: fp32 LayerNorm -> Linear (-> GEGLU) of the same
    bf16-rounded tensors, as oracle/unet.py BasicTransformerBlock computes them."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(M + C + N)
    a = _bf(torch.randn(M, C, device=cuda, generator=g))
    wp_ = _bf(torch.randn(C, C, device=cuda, generator=g) / C ** 0.5)
    res = _bf(3.0 + 2.0 * torch.randn(M, C, device=cuda, generator=g))     # rows with a large mean: the hard case
    bn = k.gemm_block_n(C, 1, 1, M)
    stats, parts = k.ln_stats_buffer(M, C, bn, cuda)
    h = k.conv_gemm(a, wp_, C, residual=res, block_n=bn, ln_stats_out=stats)           # producer
    # the partials are the row sums of the bf16 output, whatever the tiling
    assert torch.allclose(stats[..., 0].sum(1), h.float().sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[..., 1].sum(1), (h.float() ** 2).sum(1), rtol=1e-4, atol=1e-2)
    gamma = 1.0 + 0.2 * torch.randn(C, device=cuda, generator=g)
    beta = 0.3 * torch.randn(C, device=cuda, generator=g)
    w = torch.randn(N, C, device=cuda, generator=g) / C ** 0.5
    b = torch.randn(N, device=cuda, generator=g)
    side, rstd = k.ln_side(stats, C)                                                   # partials -> side tensor + rstd
    want_rstd = (h.float().var(1, unbiased=False) + 1e-5).rsqrt()
    assert torch.allclose(rstd, want_rstd, rtol=2e-3)
    assert torch.allclose(-(side[:, 0].float() + side[:, 2].float()), h.float().mean(1), rtol=1e-3, atol=1e-3)
    assert (side[:, 8:] == 0).all()
    wf = k.fold_layernorm(w, b, gamma, beta)
    assert wf.shape == (N, C + k.LN_SIDE_COLS)
    ln = F.layer_norm(h.float(), (C,), gamma, beta, 1e-5)
    ref = ln @ _bf(w).float().t() + b
    if geglu:
        bnc = k.gemm_block_n(N, 1, 1, M, k.EPI_GEGLU)
        wpk, _ = k.pack_geglu(wf, torch.zeros(N, device=cuda), bnc)
        out = k.conv_gemm(h, wpk, N, a1=side, epilogue=k.EPI_GEGLU, block_n=bnc, row_scale=rstd)   # consumer
        v, gate = ref.chunk(2, dim=-1)
        ref = v * F.gelu(gate)
    else:
        out = k.conv_gemm(h, wf, N, a1=side, row_scale=rstd)
    # what the unfused path (LayerNorm kernel -> bf16 -> GEMM) loses on the same data
    ln16 = _bf(ln)
    unf = ln16.float() @ _bf(w).float().t() + b
    if geglu:
        v, gate = unf.chunk(2, dim=-1)
        unf = v * F.gelu(gate)
    torch.cuda.synchronize()
    err, floor = _rel_err(out, ref), _rel_err(_bf(unf), ref)
    assert err < 1e-2 and err < 1.5 * floor + 2e-3, (err, floor)


@pytest.mark.parametrize("B,H,W,C,Cout", [(2, 64, 64, 320, 320), (3, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                           (1, 128, 128, 128, 128)])
def test_conv3x3_stride2_parity_views(cuda, B, H, W, C, Cout):
    """Downsample2D (3x3, stride 2, pad 1) straight from four parity views of the input -- no im2col buffer."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H + C)
    x = _bf(torch.randn(B, H, W, C, device=cuda, generator=g))
    w = _bf(torch.randn(Cout, C, 3, 3, device=cuda, generator=g) / (9 * C) ** 0.5)
    b = torch.randn(Cout, device=cuda, generator=g)
    part = k.gn_partial_buffer(B * H * W // 4, Cout, cuda)
    out = k.conv_gemm(x, k.pack_conv3x3_weight(w), Cout, taps=9, n_img=B, H=H // 2, W=W // 2, bias=b, stride=2,
                      gn_partial=part)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W // 4, Cout)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and _rel_err(out, ref) < 1e-2
    assert torch.allclose(part[..., 0].sum(0), out.float().sum(0), rtol=1e-3, atol=0.5)


@pytest.mark.parametrize("B,H,W,C,Cout", [(2, 32, 32, 640, 640), (3, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280),
                                           (1, 64, 64, 512, 512), (1, 256, 256, 128, 128)])
def test_upsample_conv_phase_form(cuda, B, H, W, C, Cout):
    """Upsample2D (nearest 2x, then 3x3) as four phase-wise 2x2 convolutions of the SOURCE with summed taps:
    no upsampled tensor, 4/9 of the multiply-adds, output written through four strided TMA views; the GroupNorm
    partials of the epilogue must still add up to the per-image channel sums."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    g = torch.Generator(device="cuda").manual_seed(H + C + 1)
    x = _bf(torch.randn(B, H, W, C, device=cuda, generator=g))
    w = _bf(torch.randn(Cout, C, 3, 3, device=cuda, generator=g) / (9 * C) ** 0.5)
    b = torch.randn(Cout, device=cuda, generator=g)
    part = k.gn_partial_buffer(B * 4 * H * W, Cout, cuda)
    out = k.conv_gemm(x, k.pack_upsample_conv_weight(w), Cout, taps=9, n_img=B, H=H, W=W, bias=b, upsample=True,
                      gn_partial=part)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(B * 4 * H * W, Cout)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and _rel_err(out, ref) < 1.2e-2
    per_img = part.view(B, -1, Cout, 2)
    got = out.float().view(B, 4 * H * W, Cout)
    assert torch.allclose(per_img[..., 0].sum(1), got.sum(1), rtol=1e-3, atol=0.5)
    assert torch.allclose(per_img[..., 1].sum(1), (got ** 2).sum(1), rtol=1e-3, atol=0.5)
