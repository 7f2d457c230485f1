"""Thin torch-tensor wrappers over the libsonic C ABI (one function per entry point).

These exist for the unit/parity tests and for host code that needs a single operator; the
sampling engine itself calls ``sonic_unet_*`` / ``sonic_latent_update`` which run whole
launch plans natively.  Tensors are only used as device-memory handles (``data_ptr``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check, lib, ptr, stream_ptr

EPI_NONE, EPI_GEGLU, EPI_QUICK_GELU = 0, 1, 2
GN_MAX_CHUNKS = 160          # SONIC_GROUPNORM_MAX_CHUNKS in include/sonic.h


def _bf16c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous(), (t.dtype, t.device, t.is_contiguous())
    return t


def _rows2d(t: torch.Tensor) -> torch.Tensor:
    """[M, N] bf16 rows with unit column stride (a column slice of a wider buffer is fine)."""
    assert t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0, \
        "expected a 2-D bf16 CUDA tensor with unit column stride and a row pitch that is a multiple of 8"
    return t


def pack_conv3x3_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """OIHW -> [tap=kh*3+kw][O][I] bf16 (K-major per tap), the layout the TMA B-box reads."""
    O, I, kh, kw = w_oihw.shape
    return w_oihw.permute(2, 3, 0, 1).reshape(kh * kw, O, I).contiguous().to(torch.bfloat16)


def pack_upsample_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """3x3 kernel applied to a nearest-2x upsampled image == four 2x2 kernels applied to the source, one per output
    phase (a, b) = (y & 1, x & 1): source row offset a - 1 + ty collects the kernel rows that land on it
    (a = 0: {0} | {1, 2};  a = 1: {0, 1} | {2}), likewise for columns.  OIHW -> [phase = 2a + b][tap = 2ty + tx][O][I]
    flattened to [16][O][I] bf16; the sums are formed in fp32 and rounded once."""
    O, I, kh, kw = w_oihw.shape
    assert (kh, kw) == (3, 3)
    w = w_oihw.float()
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = torch.zeros(4, 4, O, I, dtype=torch.float32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for ty in (0, 1):
                for tx in (0, 1):
                    acc = 0
                    for ky in rows[a][ty]:
                        for kx in rows[b][tx]:
                            acc = acc + w[:, :, ky, kx]
                    out[2 * a + b, 2 * ty + tx] = acc
    return out.reshape(16, O, I).to(torch.bfloat16).contiguous()


def pack_geglu(w: torch.Tensor, b: torch.Tensor, block_n: int):
    """Interleave the value / gate halves of ff.net.0.proj per ``block_n`` tile."""
    n2, K = w.shape
    n = n2 // 2
    half = block_n // 2
    assert n % half == 0
    wv, wg = w[:n].reshape(n // half, half, K), w[n:].reshape(n // half, half, K)
    wp = torch.cat([wv, wg], dim=1).reshape(n2, K).contiguous()
    bp = torch.cat([b[:n].reshape(-1, half), b[n:].reshape(-1, half)], dim=1).reshape(n2).contiguous()
    return wp, bp


def gemm_block_n(N: int, n_img: int, H: int, W: int, epilogue: int = EPI_NONE) -> int:
    return lib().sonic_gemm_block_n(N, n_img, H, W, epilogue)


LN_SIDE_COLS = 64            # one K chunk (128 bytes of bf16) appended to the consumer's operands


def _hi_lo(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def fold_layernorm(w, b, gamma, beta):
    """LayerNorm folded into the Linear that follows it.  Returns the consumer's weight ``[N][K + 64]`` (bf16):
    columns ``0..K-1`` = bf16(gamma .* w); columns ``K..K+7`` = (s_hi, s_lo, s_hi, s_lo, b_hi, b_lo, b_hi, b_lo) with
    s_n = sum_k of the ROUNDED gamma.w (so the mean subtraction is exact for the product the tensor cores compute)
    and b' = w beta + b; the rest zero.  With the side tensor of ``ln_side`` as the extra K chunk of A,
    ``rstd * (A W^T)`` equals ``LN(x) w^T + b``."""
    wf = w.float()
    wp = (wf * gamma.float()[None, :]).to(torch.bfloat16)
    s = wp.float().sum(dim=1)
    bp = wf @ beta.float()
    if b is not None:
        bp = bp + b.float()
    extra = torch.zeros(w.shape[0], LN_SIDE_COLS, dtype=torch.bfloat16, device=w.device)
    (s_hi, s_lo), (b_hi, b_lo) = _hi_lo(s), _hi_lo(bp)
    for col, v in enumerate((s_hi, s_lo, s_hi, s_lo, b_hi, b_lo, b_hi, b_lo)):
        extra[:, col] = v
    return torch.cat([wp, extra], dim=1).contiguous()


def ln_stats_buffer(rows, N, block_n, device):
    """Per-row (sum, sumsq) partial buffer a producer GEMM of width N / tile ``block_n`` fills; (buffer, parts)."""
    parts = 2 * ((N + block_n - 1) // block_n)
    return torch.zeros(rows, parts, 2, device=device, dtype=torch.float32), parts


def ln_side(partials, K, eps=1e-5, side=None, rstd=None):
    """Producer partials [M][parts][2] -> (side tensor [M][64] bf16, rstd [M] fp32); ``sonic_ln_side``."""
    M, parts = partials.shape[0], partials.shape[1]
    if side is None:
        side = torch.zeros(M, LN_SIDE_COLS, device=partials.device, dtype=torch.bfloat16)
    if rstd is None:
        rstd = torch.empty(M, device=partials.device, dtype=torch.float32)
    check(lib().sonic_ln_side(ptr(partials), parts, M, K, C.c_float(eps), ptr(side), ptr(rstd), stream_ptr()),
          "sonic_ln_side")
    return side, rstd


def conv_gemm(a0, w, N, *, taps=1, n_img=1, H=1, W=None, c0=None, a1=None, c1=0, bias=None,
              row_bias=None, residual=None, out=None, epilogue=EPI_NONE, block_n=0, gn_partial=None,
              ln_stats_out=None, row_scale=None, stride=1, upsample=False):
    """out[M, N'] = epilogue(implicit_gemm(A, w)); see ``sonic_conv_gemm`` in include/sonic.h.

    ``a0`` / ``a1`` are NHWC bf16 tensors whose last dim is the pixel pitch; a Linear over
    [M, K] uses the defaults (n_img=1, H=1, W=M).
    """
    _bf16c(a0), _bf16c(w)
    ld0 = a0.shape[-1]
    c0 = ld0 if c0 is None else c0
    if W is None:
        W = a0.numel() // ld0
    M = n_img * H * W * (4 if upsample else 1)      # stride 2: H, W = output extents; upsample: H, W = source extents
    n_out = N // 2 if epilogue == EPI_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a0.device, dtype=torch.bfloat16)
    args = GemmArgs()
    args.stride, args.upsample = stride, int(upsample)
    args.a0, args.c0, args.ld0 = a0.data_ptr(), c0, ld0
    if a1 is not None:
        _bf16c(a1)
        args.a1, args.c1, args.ld1 = a1.data_ptr(), c1 or a1.shape[-1], a1.shape[-1]
    args.n_img, args.H, args.W = n_img, H, W
    args.w, args.N, args.taps = w.data_ptr(), N, taps
    for name, t in (("bias", bias), ("row_bias", row_bias)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous()
            setattr(args, name, t.data_ptr())
    if residual is not None:
        _rows2d(residual)
        args.residual, args.ld_res = residual.data_ptr(), residual.stride(0)
    _rows2d(out)
    args.out, args.ld_out = out.data_ptr(), out.stride(0)
    args.epilogue, args.block_n = epilogue, block_n
    if gn_partial is not None:
        assert gn_partial.dtype == torch.float32 and gn_partial.is_contiguous() and \
            gn_partial.numel() >= (M + 31) // 32 * n_out * 2
        args.gn_partial = gn_partial.data_ptr()
    if ln_stats_out is not None:                                     # producer of a LayerNorm input
        assert block_n and ln_stats_out.dtype == torch.float32 and ln_stats_out.is_contiguous() and \
            ln_stats_out.numel() == M * 2 * ((N + block_n - 1) // block_n) * 2
        args.ln_stats_out = ln_stats_out.data_ptr()
    if row_scale is not None:                                        # folded-LayerNorm consumer: per-row rstd
        assert row_scale.dtype == torch.float32 and row_scale.numel() == M and bias is None and residual is None
        args.row_scale = row_scale.data_ptr()
    check(lib().sonic_conv_gemm(C.byref(args), stream_ptr()), "sonic_conv_gemm")
    return out


class AttentionArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
                ("ld_q", C.c_int32), ("ld_k", C.c_int32), ("ld_v", C.c_int32), ("ld_o", C.c_int32),
                ("batch", C.c_int32), ("heads", C.c_int32), ("seq_q", C.c_int32), ("seq_k", C.c_int32),
                ("head_dim", C.c_int32), ("scale", C.c_float), ("causal", C.c_int32)]


class UpdateCoeffs(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("guidance", "m_x", "m_e", "x0_x", "x0_e", "c_x", "c_e", "c_m0",
                                         "c_h1", "c_h2", "c_h3", "c_z")]


class X0Post(C.Structure):            # include/sonic.h: sonic_x0_post
    _fields_ = [("mode", C.c_int32), ("clip", C.c_float), ("p_x", C.c_float), ("p_0", C.c_float),
                ("thr", C.c_void_p), ("n_per_image", C.c_int64)]


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 0
    if t.dtype == torch.bfloat16:
        return 1
    raise TypeError(f"unsupported dtype {t.dtype}: the engine handles float32 and bfloat16 latents")


def attention_args(q, k, v, out, *, batch, heads, seq_q, seq_k, head_dim, scale=None, causal=False) -> AttentionArgs:
    for t in (q, k, v, out):
        assert t.is_cuda and t.dtype == torch.bfloat16 and t.stride(-1) == 1
    a = AttentionArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.ld_q, a.ld_k, a.ld_v, a.ld_o = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    a.batch, a.heads, a.seq_q, a.seq_k, a.head_dim = batch, heads, seq_q, seq_k, head_dim
    a.scale = float(head_dim ** -0.5 if scale is None else scale)
    a.causal = int(causal)
    return a


def attention(q, k, v, *, batch, heads, seq_q, seq_k, head_dim, scale=None, out=None, causal=False):
    """q/k/v: bf16 2-D views [batch*seq, ld] (may be column slices of a fused QKV buffer)."""
    if out is None:
        out = torch.empty((batch * seq_q, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    a = attention_args(q, k, v, out, batch=batch, heads=heads, seq_q=seq_q, seq_k=seq_k, head_dim=head_dim,
                       scale=scale, causal=causal)
    check(lib().sonic_attention(C.byref(a), stream_ptr()), "sonic_attention")
    return out


def groupnorm_scratch(n_img, groups, device):
    """Zero-initialised GroupNorm scratch: per-CTA partials, final (mean, rstd), per-image tickets
    (SONIC_GROUPNORM_SCRATCH_FLOATS in include/sonic.h); reusable across launches on one stream."""
    return torch.zeros(n_img * ((GN_MAX_CHUNKS + 1) * groups * 2 + 1), device=device, dtype=torch.float32)


def gn_partial_buffer(rows, channels, device):
    """Buffer a GEMM fills with per-32-row (sum, sumsq) of its output (``gn_partial``)."""
    return torch.zeros((rows + 31) // 32, channels, 2, device=device, dtype=torch.float32)


def groupnorm(x0, gamma, beta, *, n_img, hw, groups=32, eps=1e-5, silu=True, x1=None, out=None, part0=None,
              part1=None):
    """x0 (and optional x1): NHWC bf16 [n_img*hw, C]; returns bf16 [n_img*hw, C0+C1].  With ``part0`` (and
    ``part1``) the statistics come from the producing GEMMs' ``gn_partial`` buffers."""
    _bf16c(x0)
    c0 = x0.shape[-1]
    c1 = 0 if x1 is None else _bf16c(x1).shape[-1]
    if out is None:
        out = torch.empty((n_img * hw, c0 + c1), device=x0.device, dtype=torch.bfloat16)
    stats = groupnorm_scratch(n_img, groups, x0.device)
    if part0 is not None:
        check(lib().sonic_groupnorm_fused(ptr(x0), c0, ptr(part0), ptr(x1), c1, ptr(part1), n_img, hw, groups,
                                          C.c_float(eps), ptr(gamma), ptr(beta), int(silu), ptr(stats), ptr(out),
                                          stream_ptr()), "sonic_groupnorm_fused")
        return out
    check(lib().sonic_groupnorm_silu(ptr(x0), c0, ptr(x1), c1, n_img, hw, groups, C.c_float(eps), ptr(gamma),
                                     ptr(beta), int(silu), ptr(stats), ptr(out), stream_ptr()),
          "sonic_groupnorm_silu")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    _bf16c(x)
    rows, Cc = x.numel() // x.shape[-1], x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    check(lib().sonic_layernorm(ptr(x), ptr(out), rows, Cc, C.c_float(eps), ptr(gamma), ptr(beta), stream_ptr()),
          "sonic_layernorm")
    return out


def make_coeffs(coeffs: dict) -> UpdateCoeffs:
    k = UpdateCoeffs()
    for name, _ in UpdateCoeffs._fields_:
        setattr(k, name, float(coeffs.get(name, 0.0)))
    return k


def x0_threshold(coeffs: dict, eps, sample, *, eps_text=None, ratio=0.995, max_value=1.0):
    """Per-image dynamic-thresholding scale ``clamp(quantile(|x0|, ratio), 1, max_value)`` of the x0 prediction the
    fused update forms from (eps, sample): ``sonic_x0_threshold`` (diffusers ``_threshold_sample``, reached from
    /root/reference/src/schedulers.py:58-59,85-90).  Returns a float32 tensor [batch]."""
    k = make_coeffs(coeffs)
    for t in (eps, eps_text):
        assert t is None or (t.dtype == sample.dtype and t.is_cuda and t.is_contiguous())
    assert sample.is_cuda and sample.is_contiguous()
    n_img = sample.shape[0]
    thr = torch.empty(n_img, device=sample.device, dtype=torch.float32)
    check(lib().sonic_x0_threshold(C.byref(k), ptr(eps), ptr(eps_text), ptr(sample), n_img,
                                   C.c_int64(sample.numel() // n_img), C.c_float(ratio), C.c_float(max_value), ptr(thr),
                                   _dtype_code(sample), stream_ptr()), "sonic_x0_threshold")
    return thr


def latent_update(coeffs: dict, eps, sample, *, eps_text=None, h1=None, h2=None, h3=None, noise=None,
                  out_sample=None, out_m0=None, out_x0=None, post=None):
    """Fused CFG + scheduler update; see ``sonic_latent_update`` in include/sonic.h.  ``post``: the non-linear x0
    post-processing of ``sonic_latent_update_post`` -- ``dict(mode=1, clip=r, p_x=, p_0=)`` (clip_sample) or
    ``dict(mode=2, thr=<x0_threshold(...)>, p_x=, p_0=)`` (dynamic thresholding)."""
    k = make_coeffs(coeffs)
    code = _dtype_code(sample)
    for t in (eps, eps_text, h1, h2, h3, noise, out_sample, out_m0, out_x0):
        assert t is None or (t.dtype == sample.dtype and t.is_cuda and t.is_contiguous())
    n = sample.numel()
    n_x0 = n if out_x0 is None else min(n, out_x0.numel())      # a shorter out_x0 receives the leading images only
    if post is not None:
        thr = post.get("thr")
        assert thr is None or (thr.is_cuda and thr.dtype == torch.float32 and thr.numel() == sample.shape[0])
        q = X0Post(int(post["mode"]), float(post.get("clip", 0.0)), float(post["p_x"]), float(post["p_0"]),
                   None if thr is None else thr.data_ptr(), n // sample.shape[0])
        check(lib().sonic_latent_update_post(C.byref(k), C.byref(q), ptr(eps), ptr(eps_text), ptr(sample), ptr(h1),
                                             ptr(h2), ptr(h3), ptr(noise), ptr(out_sample), ptr(out_m0), ptr(out_x0),
                                             C.c_int64(n), C.c_int64(n_x0), code, stream_ptr()),
              "sonic_latent_update_post")
        return out_sample, out_m0, out_x0
    check(lib().sonic_latent_update_x0n(C.byref(k), ptr(eps), ptr(eps_text), ptr(sample), ptr(h1), ptr(h2), ptr(h3),
                                        ptr(noise), ptr(out_sample), ptr(out_m0), ptr(out_x0), C.c_int64(n),
                                        C.c_int64(n_x0), code, stream_ptr()), "sonic_latent_update_x0n")
    return out_sample, out_m0, out_x0


def nchw_to_nhwc8(x, dup=False):
    n, c, h, w = x.shape
    y = torch.empty((n * (2 if dup else 1), h, w, 8), device=x.device, dtype=torch.bfloat16)
    check(lib().sonic_nchw_to_nhwc8(ptr(x.contiguous()), _dtype_code(x), n, c, h * w, int(dup), ptr(y),
                                    stream_ptr()), "sonic_nchw_to_nhwc8")
    return y


def nhwc_to_nchw(x, n_img, C_, H, W, dtype):
    y = torch.empty((n_img, C_, H, W), device=x.device, dtype=dtype)
    check(lib().sonic_nhwc_to_nchw(ptr(x), x.shape[-1], n_img, C_, H * W, ptr(y), _dtype_code(y), stream_ptr()),
          "sonic_nhwc_to_nchw")
    return y


def upsample2x(x):
    n, h, w, c = x.shape
    y = torch.empty((n, 2 * h, 2 * w, c), device=x.device, dtype=torch.bfloat16)
    check(lib().sonic_upsample2x(ptr(_bf16c(x)), ptr(y), n, h, w, c, stream_ptr()), "sonic_upsample2x")
    return y


def im2col_s2(x):
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, 9 * c), device=x.device, dtype=torch.bfloat16)
    check(lib().sonic_im2col_s2(ptr(_bf16c(x)), ptr(y), n, h, w, c, stream_ptr()), "sonic_im2col_s2")
    return y


# ------------------------------------------------------------------------------------------ CLIP preprocessing
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 32 - 8 - 2
_resample_cache = {}


def _bicubic(x: float) -> float:
    """Pillow's ``bicubic_filter`` (a = -0.5), same expression order."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for a full-image resize (box = whole image):
    returns (bounds [out][2] int32, coefficients [out][ksize] int32, ksize).  Double arithmetic, C truncation."""
    import math

    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, coeffs = [], []
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        k += [0.0] * (ksize - xmax)
        bounds.append((xmin, xmax))
        coeffs.append([int(-0.5 + w * (1 << _PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << _PRECISION_BITS))
                       for w in k])
    return bounds, coeffs, ksize


def _resample_tables(H, W, size, device):
    """Resize-shortest-edge-to-``size`` + centre-crop geometry of HF ``CLIPImageProcessor`` and PIL's tables."""
    key = (H, W, size, str(device))
    if key not in _resample_cache:
        short, long_ = (H, W) if H <= W else (W, H)
        new_long = int(size * long_ / short)
        nh, nw = (size, new_long) if H <= W else (new_long, size)
        hb, hk, hks = pil_bicubic_coeffs(W, nw)
        vb, vk, vks = pil_bicubic_coeffs(H, nh)
        top, left = (nh - size) // 2, (nw - size) // 2
        max_rows = 0
        for oy0 in range(0, size, 16):
            last = min(oy0 + 16, size) - 1 + top
            max_rows = max(max_rows, vb[last][0] + vb[last][1] - vb[oy0 + top][0])
        t = lambda v: torch.tensor(v, dtype=torch.int32, device=device).contiguous()   # noqa: E731
        _resample_cache[key] = dict(hb=t(hb), hk=t(hk), hks=hks, vb=t(vb), vk=t(vk), vks=vks, top=top, left=left,
                                    max_rows=max_rows)
    return _resample_cache[key]


def clip_norm_constants(rescale_twice: bool = False):
    """(mean, std) for ``sonic_clip_preprocess``, which computes ``(v / 255 - mean) / std`` on the resized uint8 ``v``.
    ``rescale_twice`` reproduces the LITERAL behaviour of /root/reference/calc_clip_score.py:68-72 (SURVEY C-8): float
    [0,1] tensors go to the HF processor, which brings them back to [0,1] after its PIL resize and then applies its 1/255
    rescale a second time -- ``(v / 255 / 255 - mean) / std == (v / 255 - 255 mean) / (255 std)``, i.e. the same kernel
    with both constants scaled by 255 (to a grey level on natural images; HF's float path does not round / clamp the
    resized values to uint8, so bicubic overshoot on noise-like images differs by more)."""
    f = 255.0 if rescale_twice else 1.0
    return tuple(f * m for m in CLIP_MEAN), tuple(f * s for s in CLIP_STD)


def clip_preprocess(images: torch.Tensor, size: int = 224, patches_out: torch.Tensor = None, patch: int = 16,
                    rescale_twice: bool = False):
    """uint8 (n,3,H,W) -- or float / bf16 in [0,1], quantised in the kernel like base_experiment.py:198-199 -- ->
    normalised fp32 (n,3,size,size), bit-identical to HF ``CLIPImageProcessor`` on PIL; with ``patches_out`` the
    bf16 patch rows of the ViT patch-embedding GEMM are written instead (one launch, no torch ops)."""
    assert images.is_cuda and images.dim() == 4 and images.shape[1] == 3 and images.is_contiguous()
    code = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}[images.dtype]
    n, _, H, W = images.shape
    tb = _resample_tables(H, W, size, images.device)
    mean, std = clip_norm_constants(rescale_twice)
    mean, std = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    if patches_out is None:
        out, mode = torch.empty((n, 3, size, size), device=images.device, dtype=torch.float32), 0
    else:
        g = size // patch
        assert patches_out.dtype == torch.bfloat16 and patches_out.is_contiguous() and \
            tuple(patches_out.shape) == (n * g * g, 3 * patch * patch)
        out, mode = patches_out, 1
    check(lib().sonic_clip_preprocess(ptr(images), code, n, H, W, ptr(tb["hb"]), ptr(tb["hk"]), tb["hks"], ptr(tb["vb"]),
                                      ptr(tb["vk"]), tb["vks"], tb["max_rows"], tb["top"], tb["left"], size, mean, std,
                                      ptr(out), mode, patch, stream_ptr()), "sonic_clip_preprocess")
    return out
