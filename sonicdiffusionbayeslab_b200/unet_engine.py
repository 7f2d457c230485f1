"""B200 UNet engine: records SD-v1.x ``UNet2DConditionModel.forward`` into native launch plans.

This is the replacement for the ``self.unet(...)`` call of the reference's denoising loop
(/root/reference/src/models.py:227-235).  The host (this file) only *describes* the network:
it repacks a diffusers-layout state dict into the kernels' bf16 layouts, carves activation
buffers out of an arena, and records every operator into a ``sonic_plan`` (csrc/plan.cu).
Executing a forward is a single C call that replays the plan (a CUDA graph after capture):
tcgen05 implicit-GEMM convolutions / linears, tcgen05 flash attention, fused
GroupNorm+SiLU / LayerNorm / GEGLU kernels, channels-last bf16 end to end.

Plans built here:
  * ``ctx``    -- cross-attention K/V projections of the prompt embeddings (once per call;
                  they do not depend on the timestep)
  * ``full``   -- a complete forward; leaves the DeepCache feature (output of
                  up_blocks[-1].attentions[1]) resident in HBM
  * ``cached`` -- the DeepCache branch-0 step (/root/reference/src/experiments/deep_cache.py:24-29):
                  time-MLP, conv_in, last up resnet + transformer, norm_out, conv_out
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import kernels as K
from ._lib import GemmArgs, check, lib


@dataclass
class UNetArch:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attn_blocks: tuple = (True, True, True, False)
    num_heads: int = 8
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5


def deepcache_runs(branch: int, n_blocks: int, layers: int):
    """Which layers a DeepCache non-refresh step recomputes for ``cache_branch_id = branch`` (SURVEY appendix A.4,
    ``DeepCacheSDHelper.is_skip_step``): returns ``(down_runs(b, j), up_runs(b, j), first_up)`` with b / j in
    FORWARD order (a down-sampler is layer ``layers``); ``first_up`` = the (up block, layer) whose input is the
    cached feature."""
    bid, lid = divmod(branch, 3)

    def down_runs(b, j):
        return b < bid or (b == bid and j < lid)

    def up_runs(b, j):
        bi, li = n_blocks - 1 - b, layers - j               # DeepCache indexes the up path in reverse
        return bi < bid or (bi == bid and li <= lid)

    return down_runs, up_runs, (n_blocks - 1 - bid, layers - lid)


class Arena:
    """Exact-size free lists over torch-owned device memory (pointers stay valid for the plan)."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.all = []
        self.bytes = 0

    def alloc(self, shape, dtype=torch.bfloat16):
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        nbytes = (nbytes + 1023) // 1024 * 1024
        lst = self.free.get(nbytes)
        if lst:
            raw = lst.pop()
        else:
            raw = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.all.append(raw)
            self.bytes += nbytes
        t = raw[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        t._arena_raw = raw
        return t

    def release(self, t):
        for attr in ("_gn_part", "_ln_part"):     # GroupNorm / LayerNorm pre-reduction buffers travel with their tensor
            part = getattr(t, attr, None)
            if part is not None:
                setattr(t, attr, None)
                self.release(part)
        raw = t._arena_raw
        self.free.setdefault(raw.numel(), []).append(raw)


class _Plan:
    def __init__(self):
        self.h = C.c_void_p()
        self.log = []                      # one human-readable line per recorded operator, in order
        check(lib().sonic_plan_create(C.byref(self.h)), "sonic_plan_create")

    def run(self, stream_ptr):
        check(lib().sonic_plan_run(self.h, stream_ptr), "sonic_plan_run")

    def capture(self, stream_ptr):
        check(lib().sonic_plan_capture(self.h, stream_ptr), "sonic_plan_capture")

    def stats(self):
        n, f = C.c_int32(), C.c_double()
        check(lib().sonic_plan_stats(self.h, C.byref(n), C.byref(f)), "sonic_plan_stats")
        return n.value, f.value

    def profile(self, stream_ptr, max_ops=4096):
        """Per-operator device time of one eager run: list of (kind, ms, flops)."""
        ms = (C.c_float * max_ops)()
        kinds = (C.c_int32 * max_ops)()
        flops = (C.c_double * max_ops)()
        n = C.c_int32()
        check(lib().sonic_plan_profile(self.h, stream_ptr, max_ops, ms, kinds, flops, C.byref(n)),
              "sonic_plan_profile")
        return [(kinds[i], ms[i], flops[i]) for i in range(n.value)]

    def __del__(self):
        try:
            if self.h:
                lib().sonic_plan_destroy(self.h)
        except Exception:
            pass


class PackedWeights:
    """A diffusers-layout UNet state dict plus the cache of tensors repacked for the kernels
    (bf16 K-major matrices, [tap][O][I] convolutions, fused QKV / KV, tile-interleaved GEGLU).
    Shared by every ``UNetEngine`` specialisation of one model so weights live in HBM once."""

    def __init__(self, state_dict, device="cuda"):
        self.sd = state_dict
        self.device = torch.device(device)
        self.cache = {}

    def nbytes(self):
        tot = 0
        for v in self.cache.values():
            for t in (v if isinstance(v, tuple) else (v,)):
                if isinstance(t, torch.Tensor):
                    tot += t.numel() * t.element_size()
        return tot


class UNetEngine:
    """One engine = one (effective batch, resolution, io dtype) specialisation of the UNet."""

    def __init__(self, weights, *, n_latents: int, cfg_dup: bool, arch: UNetArch = UNetArch(),
                 height: int = 64, width: int = 64, ctx_len: int = 77, io_dtype=torch.bfloat16,
                 device="cuda", build_cached: bool = True, cache_branch: int = 0):
        if not isinstance(weights, PackedWeights):
            weights = PackedWeights(weights, device)
        state_dict = weights.sd
        self.weights = weights
        self.arch = arch
        self.dev = torch.device(device)
        self.n_lat = n_latents
        self.n = n_latents * (2 if cfg_dup else 1)          # UNet batch
        if not 0 <= cache_branch < 3 * len(arch.block_out_channels):
            raise ValueError(f"cache_branch_id {cache_branch} out of range for {len(arch.block_out_channels)} blocks")
        self.cache_branch = cache_branch                     # DeepCache branch the ``cached`` plan is recorded for
        self.cfg_dup = cfg_dup
        self.H, self.W, self.ctx_len = height, width, ctx_len
        self.io_dtype = io_dtype
        self.arena = Arena(self.dev)
        self.fuse_gn_stats = True                            # GroupNorm statistics from the producers' epilogues
        self.fold_layernorm = True                           # LayerNorms folded into the GEMMs around them
        self._keep = []                                      # packed weights
        self.sd = state_dict
        a = arch
        # fixed I/O buffers the plans read / write
        self.x_in = torch.zeros(n_latents, a.in_channels, height, width, device=self.dev, dtype=io_dtype)
        self.eps = torch.zeros(self.n, a.out_channels, height, width, device=self.dev, dtype=io_dtype)
        self.t_dev = torch.zeros(1, device=self.dev, dtype=torch.float32)
        self.ctx = torch.zeros(self.n * ctx_len, a.cross_attention_dim, device=self.dev, dtype=torch.bfloat16)
        self._w = weights.cache
        self._temb_jobs = []
        self._ctx_kv = {}
        self._ln_bufs = {}
        self.cache_feature = None
        self.plans = {}
        with torch.no_grad():
            self.plans["ctx"] = self._build_ctx_plan()
            self.plans["full"] = self._build_unet_plan(cached=False)
            if build_cached:
                self.plans["cached"] = self._build_unet_plan(cached=True)
        self._graph_stream = None

    # ------------------------------------------------------------------ weights
    def _p(self, name):
        return self.sd[name].detach().to(self.dev)

    def _src(self, name):
        """The checkpoint tensor WHERE IT LIVES: repacking (cast / permute / cat) runs on the source device and only
        the packed result is copied to HBM -- a CPU-resident state dict costs no device launches to pack."""
        return self.sd[name].detach()

    def _f32(self, name):
        key = ("f32", name)
        if key not in self._w:
            self._w[key] = self._src(name).float().contiguous().to(self.dev)
        return self._w[key]

    def _lin(self, name):
        key = ("lin", name)
        if key not in self._w:
            w = self._src(name)
            if w.dim() == 4:                                  # 1x1 conv
                w = w.reshape(w.shape[0], w.shape[1])
            self._w[key] = w.to(torch.bfloat16).contiguous().to(self.dev)
        return self._w[key]

    def _folded(self, key, weight_names, bias_name, norm_prefix):
        """``K.fold_layernorm`` weight [N][K + 64] for the (row-concatenated) Linear behind LayerNorm ``norm_prefix``."""
        if key + ("ln",) not in self._w:
            w = torch.cat([self._src(n) for n in weight_names], 0)
            b = self._src(bias_name) if bias_name else None
            self._w[key + ("ln",)] = K.fold_layernorm(w, b, self._src(norm_prefix + ".weight"),
                                                      self._src(norm_prefix + ".bias")).to(self.dev)
        return self._w[key + ("ln",)]

    def _ln_fold(self, plan, h, C_ln):
        """Between producer and consumer: per-row partials of ``h`` -> (side tensor, rstd).  The buffers are shared by
        every LayerNorm of that row count (plans run in stream order); side columns 8..63 stay zero for ever."""
        M = h.shape[0]
        if M not in self._ln_bufs:
            self._ln_bufs[M] = (torch.zeros(M, K.LN_SIDE_COLS, device=self.dev, dtype=torch.bfloat16),
                                torch.empty(M, device=self.dev, dtype=torch.float32))
        side, rstd = self._ln_bufs[M]
        part = h._ln_part
        check(lib().sonic_plan_add_ln_side(plan.h, K.ptr(part), part.shape[1], M, C_ln, C.c_float(1e-5), K.ptr(side),
                                           K.ptr(rstd)), "sonic_plan_add_ln_side")
        plan.log.append(f"ln_side rows={M} C={C_ln} parts={part.shape[1]}")
        return side, rstd

    def _conv3(self, name):
        key = ("c3", name)
        if key not in self._w:
            self._w[key] = K.pack_conv3x3_weight(self._src(name)).to(self.dev)
        return self._w[key]

    # ------------------------------------------------------------------ op recording helpers
    def _gemm(self, plan, a0, w, N, *, n_img=1, H=1, W=None, taps=1, c0=None, a1=None, bias=None,
              residual=None, out=None, epilogue=K.EPI_NONE, block_n=0, gn_stats=False, ln_stats=False, row_scale=None,
              stride=1, upsample=False):
        ld0 = a0.shape[-1]
        M = a0.numel() // ld0
        if stride == 2:                                       # H, W are the output extents; A is the 2H x 2W input
            M //= 4
        elif upsample:                                        # H, W are the source extents; the output is 2H x 2W
            M *= 4
        if W is None:
            W = M
        n_out = N // 2 if epilogue == K.EPI_GEGLU else N
        if out is None:
            out = self.arena.alloc((M, n_out))
        g = GemmArgs()
        g.a0, g.c0, g.ld0 = a0.data_ptr(), (ld0 if c0 is None else c0), ld0
        if a1 is not None:
            g.a1, g.c1, g.ld1 = a1.data_ptr(), a1.shape[-1], a1.shape[-1]
        g.n_img, g.H, g.W = n_img, H, W
        g.w, g.N, g.taps = w.data_ptr(), N, taps
        if bias is not None:
            assert bias.dtype == torch.float32
            g.bias = bias.data_ptr()
        if residual is not None:
            g.residual, g.ld_res = residual.data_ptr(), residual.shape[-1]
        g.out, g.ld_out = out.data_ptr(), out.shape[-1]
        g.epilogue, g.block_n = epilogue, block_n
        g.stride, g.upsample = stride, int(upsample)
        if ln_stats:
            # producer of a LayerNorm input: per-row (sum, sumsq) partials of the output ride along with the tensor
            bn = block_n or K.gemm_block_n(N, n_img, H, W, epilogue)
            g.block_n = bn
            parts = 2 * ((N + bn - 1) // bn)
            buf = getattr(out, "_ln_part", None)
            if buf is None or buf.shape[1] != parts:
                buf = self.arena.alloc((M, parts, 2), torch.float32)
                out._ln_part = buf
            g.ln_stats_out = buf.data_ptr()
        if row_scale is not None:                            # folded-LayerNorm consumer (a1 = the side tensor)
            assert row_scale.numel() == M and bias is None and residual is None
            g.row_scale = row_scale.data_ptr()
        if gn_stats and self.fuse_gn_stats:
            # the epilogue also writes per-32-row (sum, sumsq) of the output: the consumer GroupNorm needs no
            # statistics pass over the tensor
            part = getattr(out, "_gn_part", None)
            if part is None:
                part = self.arena.alloc(((M + 31) // 32, n_out, 2), torch.float32)
                out._gn_part = part
            g.gn_partial = part.data_ptr()
        check(lib().sonic_plan_add_conv_gemm(plan.h, C.byref(g)), "sonic_plan_add_conv_gemm")
        kk = g.c0 + (g.c1 if a1 is not None else 0)
        plan.log.append(f"gemm M={M} N={N} K={kk}x{taps}{'s2' if stride == 2 else 'up' if upsample else ''} "
                        f"img={n_img}x{H}x{W} epi={epilogue}"
                        f"{' +res' if residual is not None else ''}{' +lnfold' if row_scale is not None else ''}"
                        f"{' +lnstats' if ln_stats else ''}")
        return out

    def _gn(self, plan, x0, x1, prefix, hw, eps, silu):
        c = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
        y = self.arena.alloc((x0.shape[0], c))
        if not hasattr(self, "_gn_stats"):                    # one scratch: plans run in stream order
            self._gn_stats = K.groupnorm_scratch(self.n, self.arch.norm_num_groups, self.dev)
        stats = self._gn_stats
        p0 = getattr(x0, "_gn_part", None)
        p1 = None if x1 is None else getattr(x1, "_gn_part", None)
        if self.fuse_gn_stats and p0 is not None and (x1 is None or p1 is not None) and hw % 32 == 0:
            check(lib().sonic_plan_add_groupnorm_fused(
                plan.h, K.ptr(x0), x0.shape[-1], K.ptr(p0), K.ptr(x1), 0 if x1 is None else x1.shape[-1], K.ptr(p1),
                self.n, hw, self.arch.norm_num_groups, C.c_float(eps), K.ptr(self._f32(prefix + ".weight")),
                K.ptr(self._f32(prefix + ".bias")), int(silu), K.ptr(stats), K.ptr(y)),
                "sonic_plan_add_groupnorm_fused")
            plan.log.append(f"groupnorm rows={x0.shape[0]} C={c} silu={int(silu)} fused-stats")
            return y
        check(lib().sonic_plan_add_groupnorm(
            plan.h, K.ptr(x0), x0.shape[-1], K.ptr(x1), 0 if x1 is None else x1.shape[-1], self.n, hw,
            self.arch.norm_num_groups, C.c_float(eps), K.ptr(self._f32(prefix + ".weight")),
            K.ptr(self._f32(prefix + ".bias")), int(silu), K.ptr(stats), K.ptr(y)), "sonic_plan_add_groupnorm")
        plan.log.append(f"groupnorm rows={x0.shape[0]} C={c} silu={int(silu)}")
        return y

    def _ln(self, plan, x, prefix):
        y = self.arena.alloc(tuple(x.shape))
        check(lib().sonic_plan_add_layernorm(plan.h, K.ptr(x), K.ptr(y), x.shape[0], x.shape[1], C.c_float(1e-5),
                                             K.ptr(self._f32(prefix + ".weight")),
                                             K.ptr(self._f32(prefix + ".bias"))), "sonic_plan_add_layernorm")
        plan.log.append(f"layernorm rows={x.shape[0]} C={x.shape[1]}")
        return y

    def _attn(self, plan, q, k, v, seq_q, seq_k, d):
        out = self.arena.alloc((self.n * seq_q, self.arch.num_heads * d))
        a = K.attention_args(q, k, v, out, batch=self.n, heads=self.arch.num_heads, seq_q=seq_q, seq_k=seq_k,
                             head_dim=d)
        check(lib().sonic_plan_add_attention(plan.h, C.byref(a)), "sonic_plan_add_attention")
        plan.log.append(f"attention B={self.n} H={self.arch.num_heads} Sq={seq_q} Sk={seq_k} d={d}")
        return out

    # ------------------------------------------------------------------ blocks
    def _resnet(self, plan, prefix, x0, x1, H, W, cout):
        """ResnetBlock2D over the channel concat (x0 | x1); returns a fresh [M, cout] buffer."""
        hw = H * W
        cin = x0.shape[-1] + (0 if x1 is None else x1.shape[-1])
        h = self._gn(plan, x0, x1, prefix + ".norm1", hw, self.arch.norm_eps, True)
        tb = self._temb_bias[prefix]                          # conv1.bias + time_emb_proj(silu(temb))
        h1 = self._gemm(plan, h, self._conv3(prefix + ".conv1.weight"), cout, n_img=self.n, H=H, W=W, taps=9,
                        bias=tb, gn_stats=True)
        self.arena.release(h)
        h2 = self._gn(plan, h1, None, prefix + ".norm2", hw, self.arch.norm_eps, True)
        self.arena.release(h1)
        sc = None
        if cin != cout:
            sc = self._gemm(plan, x0, self._lin(prefix + ".conv_shortcut.weight"), cout, n_img=self.n, H=H, W=W,
                            a1=x1, bias=self._f32(prefix + ".conv_shortcut.bias"))
            res = sc
        else:
            assert x1 is None
            res = x0
        out = self._gemm(plan, h2, self._conv3(prefix + ".conv2.weight"), cout, n_img=self.n, H=H, W=W, taps=9,
                         bias=self._f32(prefix + ".conv2.bias"), residual=res, gn_stats=True)
        self.arena.release(h2)
        if sc is not None:
            self.arena.release(sc)
        return out

    def _transformer(self, plan, prefix, x, H, W):
        """Transformer2DModel (depth 1); the result overwrites ``x`` (residual fused in place)."""
        a = self.arch
        hw = H * W
        Cc = x.shape[-1]
        d = Cc // a.num_heads
        tb = prefix + ".transformer_blocks.0"
        g = self._gn(plan, x, None, prefix + ".norm", hw, 1e-6, False)
        fold = self.fold_layernorm
        h = self._gemm(plan, g, self._lin(prefix + ".proj_in.weight"), Cc, bias=self._f32(prefix + ".proj_in.bias"),
                       ln_stats=fold)
        self.arena.release(g)
        # The three LayerNorms of the block launch nothing: the producer of ``h`` leaves per-row (sum, sumsq) in its
        # epilogue and the consuming projection applies rstd / mean after the product (``K.fold_layernorm``).
        # self-attention: fused QKV projection, heads read in place by the attention kernel
        if fold:
            side, rstd = self._ln_fold(plan, h, Cc)
            wq = self._folded(("qkv", tb), [tb + f".attn1.to_{n}.weight" for n in "qkv"], None, tb + ".norm1")
            qkv = self._gemm(plan, h, wq, 3 * Cc, a1=side, row_scale=rstd)
        else:
            ln = self._ln(plan, h, tb + ".norm1")
            key = ("qkv", tb)
            if key not in self._w:
                self._w[key] = torch.cat([self._src(tb + f".attn1.to_{n}.weight") for n in "qkv"], 0) \
                    .to(torch.bfloat16).contiguous().to(self.dev)
            qkv = self._gemm(plan, ln, self._w[key], 3 * Cc)
            self.arena.release(ln)
        ao = self._attn(plan, qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], hw, hw, d)
        self.arena.release(qkv)
        self._gemm(plan, ao, self._lin(tb + ".attn1.to_out.0.weight"), Cc, bias=self._f32(tb + ".attn1.to_out.0.bias"),
                   residual=h, out=h, ln_stats=fold)
        self.arena.release(ao)
        # cross-attention: K/V come from the per-call ctx plan
        if fold:
            side, rstd = self._ln_fold(plan, h, Cc)
            wq = self._folded(("xq", tb), [tb + ".attn2.to_q.weight"], None, tb + ".norm2")
            q = self._gemm(plan, h, wq, Cc, a1=side, row_scale=rstd)
        else:
            ln = self._ln(plan, h, tb + ".norm2")
            q = self._gemm(plan, ln, self._lin(tb + ".attn2.to_q.weight"), Cc)
            self.arena.release(ln)
        kv = self._ctx_kv[tb]
        ao = self._attn(plan, q, kv[:, :Cc], kv[:, Cc:], hw, self.ctx_len, d)
        self.arena.release(q)
        self._gemm(plan, ao, self._lin(tb + ".attn2.to_out.0.weight"), Cc, bias=self._f32(tb + ".attn2.to_out.0.bias"),
                   residual=h, out=h, ln_stats=fold)
        self.arena.release(ao)
        # feed-forward: GEGLU fused into the first GEMM's epilogue
        bn = K.gemm_block_n(8 * Cc, 1, 1, h.shape[0], K.EPI_GEGLU)
        if fold:
            key = ("geglu_ln", tb, bn)
            if key not in self._w:
                wf = K.fold_layernorm(self._src(tb + ".ff.net.0.proj.weight"), self._src(tb + ".ff.net.0.proj.bias"),
                                      self._src(tb + ".norm3.weight"), self._src(tb + ".norm3.bias"))
                wp, _ = K.pack_geglu(wf, torch.zeros(wf.shape[0], device=wf.device), bn)       # rows interleaved per tile; no bias vector
                self._w[key] = wp.to(self.dev)
            side, rstd = self._ln_fold(plan, h, Cc)
            ff = self._gemm(plan, h, self._w[key], 8 * Cc, a1=side, epilogue=K.EPI_GEGLU, block_n=bn, row_scale=rstd)
        else:
            ln = self._ln(plan, h, tb + ".norm3")
            key = ("geglu", tb, bn)
            if key not in self._w:
                wp, bp = K.pack_geglu(self._src(tb + ".ff.net.0.proj.weight").to(torch.bfloat16),
                                      self._src(tb + ".ff.net.0.proj.bias").float(), bn)
                self._w[key] = (wp.to(self.dev), bp.to(self.dev), bn)
            wp, bp, bn = self._w[key]
            ff = self._gemm(plan, ln, wp, 8 * Cc, bias=bp, epilogue=K.EPI_GEGLU, block_n=bn)
            self.arena.release(ln)
        self._gemm(plan, ff, self._lin(tb + ".ff.net.2.weight"), Cc, bias=self._f32(tb + ".ff.net.2.bias"),
                   residual=h, out=h)
        self.arena.release(ff)
        self._gemm(plan, h, self._lin(prefix + ".proj_out.weight"), Cc, bias=self._f32(prefix + ".proj_out.bias"),
                   residual=x, out=x, gn_stats=True)
        self.arena.release(h)
        return x

    # ------------------------------------------------------------------ plans
    def _attn_prefixes(self):
        a = self.arch
        out = []
        for b, has in enumerate(a.attn_blocks):
            if has:
                out += [f"down_blocks.{b}.attentions.{j}" for j in range(a.layers_per_block)]
        out.append("mid_block.attentions.0")
        for b, has in enumerate(reversed(a.attn_blocks)):
            if has:
                out += [f"up_blocks.{b}.attentions.{j}" for j in range(a.layers_per_block + 1)]
        return out

    def _resnet_prefixes(self):
        a = self.arch
        out = []
        for b in range(len(a.block_out_channels)):
            out += [f"down_blocks.{b}.resnets.{j}" for j in range(a.layers_per_block)]
        out += ["mid_block.resnets.0", "mid_block.resnets.1"]
        for b in range(len(a.block_out_channels)):
            out += [f"up_blocks.{b}.resnets.{j}" for j in range(a.layers_per_block + 1)]
        return out

    def _downsample(self, plan, b, h, H, W, cout):
        """Stride-2 3x3 convolution of ``down_blocks[b].downsamplers[0]``: nine taps over four parity views of the
        input straight from the TMA unit (no im2col buffer); returns (out, H/2, W/2)."""
        out = self._gemm(plan, h, self._conv3(f"down_blocks.{b}.downsamplers.0.conv.weight"), cout, n_img=self.n,
                         H=H // 2, W=W // 2, taps=9, stride=2,
                         bias=self._f32(f"down_blocks.{b}.downsamplers.0.conv.bias"), gn_stats=True)
        return out, H // 2, W // 2

    def _up3(self, name):
        key = ("up3", name)
        if key not in self._w:
            self._w[key] = K.pack_upsample_conv_weight(self._src(name)).to(self.dev)
        return self._w[key]

    def _upsample_conv(self, plan, prefix, h, H, W, cout):
        """``Upsample2D``: nearest 2x + 3x3 convolution as four phase-wise 2x2 convolutions of the SOURCE (4/9 of the
        multiply-adds, no upsampled tensor); returns (out, 2H, 2W)."""
        out = self._gemm(plan, h, self._up3(prefix + ".conv.weight"), cout, n_img=self.n, H=H, W=W, taps=9,
                         upsample=True, bias=self._f32(prefix + ".conv.bias"), gn_stats=True)
        return out, 2 * H, 2 * W

    def _build_ctx_plan(self):
        plan = _Plan()
        for pfx in self._attn_prefixes():
            tb = pfx + ".transformer_blocks.0"
            if ("kv", tb) not in self._w:
                wk, wv = self._src(tb + ".attn2.to_k.weight"), self._src(tb + ".attn2.to_v.weight")
                self._w[("kv", tb)] = torch.cat([wk, wv], 0).to(torch.bfloat16).contiguous().to(self.dev)
            w = self._w[("kv", tb)]
            kv = self.arena.alloc((self.n * self.ctx_len, w.shape[0]))
            self._ctx_kv[tb] = kv                              # persistent: never released
            self._gemm(plan, self.ctx, w, w.shape[0], out=kv)
        return plan

    def _record_time_path(self, plan):
        """sinusoid -> time MLP -> every resnet's time_emb_proj, folded with conv1.bias (M=1 GEMVs)."""
        a = self.arch
        c0 = a.block_out_channels[0]
        T = 4 * c0
        if not hasattr(self, "_temb_bias"):
            self._t_sin = torch.zeros(c0, device=self.dev, dtype=torch.float32)
            self._t_e1 = torch.zeros(T, device=self.dev, dtype=torch.float32)
            self._t_e2 = torch.zeros(T, device=self.dev, dtype=torch.float32)
            self._temb_bias = {}
            for pfx in self._resnet_prefixes():
                n = self.sd[pfx + ".time_emb_proj.weight"].shape[0]
                self._temb_bias[pfx] = torch.zeros(n, device=self.dev, dtype=torch.float32)
        check(lib().sonic_plan_add_timestep_embedding(plan.h, K.ptr(self.t_dev), c0, K.ptr(self._t_sin)),
              "sonic_plan_add_timestep_embedding")
        plan.log.append("timestep_embedding")

        def gemv(jobs, x, Kdim, silu_in):
            n = len(jobs)
            arr_w = (C.c_void_p * n)(*[j[0].data_ptr() for j in jobs])
            arr_b = (C.c_void_p * n)(*[j[1].data_ptr() for j in jobs])
            arr_a = (C.c_void_p * n)(*[0 if j[2] is None else j[2].data_ptr() for j in jobs])
            arr_y = (C.c_void_p * n)(*[j[3].data_ptr() for j in jobs])
            arr_n = (C.c_int32 * n)(*[j[3].numel() for j in jobs])
            check(lib().sonic_plan_add_gemv(plan.h, n, arr_w, arr_b, arr_a, arr_y, arr_n, K.ptr(x), Kdim,
                                            int(silu_in)), "sonic_plan_add_gemv")
            plan.log.append(f"gemv jobs={n} rows={sum(j[3].numel() for j in jobs)} K={Kdim}")

        gemv([(self._lin("time_embedding.linear_1.weight"), self._f32("time_embedding.linear_1.bias"), None,
               self._t_e1)], self._t_sin, c0, False)
        gemv([(self._lin("time_embedding.linear_2.weight"), self._f32("time_embedding.linear_2.bias"), None,
               self._t_e2)], self._t_e1, T, True)
        return gemv

    def _build_unet_plan(self, cached: bool):
        """One walk over the network serves both plans.  ``cached`` = a DeepCache non-refresh step for
        ``self.cache_branch`` (appendix A.4: ``block_id, layer_id = divmod(branch, 3)``): down layer (b, j) runs iff
        ``b < block_id or (b == block_id and j < layer_id)`` (a down-sampler is layer ``layers_per_block``), the mid
        block never, up layer (b, j) -- DeepCache indexes the up path in reverse, ``bi = nb-1-b``,
        ``li = layers-1-j`` -- iff ``bi < block_id or (bi == block_id and li <= layer_id)``; the first up layer that
        runs takes as input the feature the last FULL step left resident in HBM (``self.cache_feature``).  The skip
        connections a cached step consumes are exactly the ones it recomputes."""
        a = self.arch
        plan = _Plan()
        n, H, W = self.n, self.H, self.W
        boc = a.block_out_channels
        nb, L = len(boc), a.layers_per_block
        d_runs, u_runs, first_up = deepcache_runs(self.cache_branch, nb, L)

        def down_runs(b, j):
            return not cached or d_runs(b, j)

        def up_runs(b, j):
            return not cached or u_runs(b, j)

        gemv = self._record_time_path(plan)
        prefixes = []
        for b in range(nb):
            prefixes += [f"down_blocks.{b}.resnets.{j}" for j in range(L) if down_runs(b, j)]
        if not cached:
            prefixes += ["mid_block.resnets.0", "mid_block.resnets.1"]
        for b in range(nb):
            prefixes += [f"up_blocks.{b}.resnets.{j}" for j in range(L + 1) if up_runs(b, j)]
        gemv([(self._lin(p + ".time_emb_proj.weight"), self._f32(p + ".time_emb_proj.bias"),
               self._f32(p + ".conv1.bias"), self._temb_bias[p]) for p in prefixes], self._t_e2, 4 * boc[0], True)

        # conv_in: NCHW latents -> NHWC (8 channels, zero padded, CFG duplication) -> tcgen05 conv
        x8 = self.arena.alloc((n * H * W, 8))
        check(lib().sonic_plan_add_nchw_to_nhwc8(plan.h, K.ptr(self.x_in), K._dtype_code(self.x_in), self.n_lat,
                                                 a.in_channels, H * W, int(self.cfg_dup), K.ptr(x8)),
              "sonic_plan_add_nchw_to_nhwc8")
        plan.log.append("nchw_to_nhwc8")
        # 3x3 patches of the 8-channel input as ONE 144-byte row per pixel, then a K = 72 GEMM: as nine shifted
        # TMA taps the A operand arrives in 16-byte rows and the copy engine's row rate made this 154 us (38 TFLOP/s).
        key = ("conv_in",)
        if key not in self._w:
            w = self._src("conv_in.weight")
            wp = torch.zeros(w.shape[0], 8, 3, 3, device=w.device, dtype=w.dtype)
            wp[:, : w.shape[1]] = w
            self._w[key] = wp.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous().to(self.dev)
        col = self.arena.alloc((n * H * W, 72))
        check(lib().sonic_plan_add_im2col3x3(plan.h, K.ptr(x8), K.ptr(col), n, H, W, 8, 1), "sonic_plan_add_im2col3x3")
        plan.log.append(f"im2col3x3 {n}x{H}x{W}x8")
        self.arena.release(x8)
        h = self._gemm(plan, col, self._w[key], boc[0], bias=self._f32("conv_in.bias"), gn_stats=True)
        self.arena.release(col)

        if cached:
            assert self.cache_feature is not None, "build the full plan first"
        # ---- down path; ``skips`` holds (tensor or None for a layer this plan does not run, H, W)
        skips = [(h, H, W)]
        alive = True                                        # False once the walk passes the last layer that runs
        for b, cout in enumerate(boc):
            for j in range(L):
                if alive and down_runs(b, j):
                    r = self._resnet(plan, f"down_blocks.{b}.resnets.{j}", h, None, H, W, cout)
                    if a.attn_blocks[b]:
                        r = self._transformer(plan, f"down_blocks.{b}.attentions.{j}", r, H, W)
                    h = r
                    skips.append((h, H, W))
                else:
                    alive = False
                    skips.append((None, H, W))
            if b != nb - 1:
                if alive and down_runs(b, L):
                    h, H, W = self._downsample(plan, b, h, H, W, cout)
                    skips.append((h, H, W))
                else:
                    alive = False
                    H, W = H // 2, W // 2
                    skips.append((None, H, W))
        # ---- mid
        if not cached:
            c = boc[-1]
            r = self._resnet(plan, "mid_block.resnets.0", h, None, H, W, c)
            r = self._transformer(plan, "mid_block.attentions.0", r, H, W)
            h = self._resnet(plan, "mid_block.resnets.1", r, None, H, W, c)
            self.arena.release(r)
        else:
            h = None
        # ---- up path
        rev = list(reversed(boc))
        rev_attn = list(reversed(a.attn_blocks))
        for b, cout in enumerate(rev):
            for j in range(L + 1):
                skip, sh, sw = skips.pop()
                if (b, j) == first_up:
                    if cached:
                        h = self.cache_feature              # left resident by the last full step
                    else:
                        self.cache_feature = h              # never released: stays in HBM between steps
                if not up_runs(b, j):
                    assert skip is None, "a cached step recomputed a skip connection it does not consume"
                    continue
                assert (sh, sw) == (H, W) and skip is not None and h is not None
                r = self._resnet(plan, f"up_blocks.{b}.resnets.{j}", h, skip, H, W, cout)
                if h is not self.cache_feature:
                    self.arena.release(h)
                self.arena.release(skip)
                if rev_attn[b]:
                    r = self._transformer(plan, f"up_blocks.{b}.attentions.{j}", r, H, W)
                h = r
            if b != nb - 1:
                if h is None:                               # cached plan: this block did not run at all
                    H, W = 2 * H, 2 * W
                    continue
                src = h
                h, H, W = self._upsample_conv(plan, f"up_blocks.{b}.upsamplers.0", src, H, W, cout)
                if src is not self.cache_feature:
                    self.arena.release(src)
        assert not skips
        # ---- out: GroupNorm+SiLU -> conv3x3 (4 output channels padded to one 16-wide MMA tile) -> NCHW
        g = self._gn(plan, h, None, "conv_norm_out", self.H * self.W, a.norm_eps, True)
        self.arena.release(h)
        key = ("conv_out",)
        if key not in self._w:
            w = self._src("conv_out.weight")
            wp = torch.zeros(16, w.shape[1], 3, 3, device=w.device, dtype=w.dtype)
            wp[: w.shape[0]] = w
            bp = torch.zeros(16, device=w.device, dtype=torch.float32)
            bp[: w.shape[0]] = self._src("conv_out.bias").float()
            self._w[key] = (K.pack_conv3x3_weight(wp).to(self.dev), bp.to(self.dev))
        wp, bp = self._w[key]
        o16 = self._gemm(plan, g, wp, 16, n_img=n, H=self.H, W=self.W, taps=9, bias=bp, block_n=16)
        self.arena.release(g)
        check(lib().sonic_plan_add_nhwc_to_nchw(plan.h, K.ptr(o16), 16, n, a.out_channels, self.H * self.W,
                                                K.ptr(self.eps), K._dtype_code(self.eps)),
              "sonic_plan_add_nhwc_to_nchw")
        plan.log.append("nhwc_to_nchw")
        self.arena.release(o16)
        return plan

    # ------------------------------------------------------------------ execution
    def set_context(self, prompt_embeds: torch.Tensor):
        """prompt_embeds: [n, ctx_len, cross_dim]; projects K/V of all cross-attention layers."""
        assert prompt_embeds.shape == (self.n, self.ctx_len, self.arch.cross_attention_dim), prompt_embeds.shape
        self.ctx.copy_(prompt_embeds.reshape(self.n * self.ctx_len, -1))
        self.plans["ctx"].run(K.stream_ptr())

    def capture_graphs(self):
        """Instantiate one CUDA graph per plan (replayed by ``forward``)."""
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for name, p in self.plans.items():
                if name != "ctx":
                    p.capture(C.c_void_p(s.cuda_stream))
        torch.cuda.current_stream().wait_stream(s)
        self._graph_stream = s

    def forward(self, t: float, cached: bool = False) -> torch.Tensor:
        """Runs on the current stream: reads ``x_in`` / the projected context, writes ``eps``."""
        self.t_dev.fill_(float(t))
        self.plans["cached" if cached else "full"].run(K.stream_ptr())
        return self.eps

    def stats(self, name="full"):
        return self.plans[name].stats()
