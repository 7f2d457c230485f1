"""Oracle restatement of DeepCache 0.1.1 ``DeepCacheSDHelper`` (test infrastructure only).

The reference drives it at /root/reference/src/experiments/deep_cache.py:24-29,58
(``set_params(cache_interval, cache_branch_id)``, ``enable()``, ``disable()``); the mechanism
itself lives in the DeepCache package (poetry.lock:436-437), absent here, so its published
wrapper semantics are restated over the oracle UNet:

  * every down/up block, resnet, attention, down/up-sampler and the mid block is wrapped; a
    wrapped call returns its cached output when ``is_skip_step`` says so, else computes+stores;
  * ``is_skip_step``: full step iff ``(cur - start) % interval == 0``; otherwise skip when
    ``block_i > cache_block_id`` or mid; compute when ``block_i < cache_block_id``; at the
    boundary block skip iff ``layer_i >= cache_layer_id`` (down) / ``layer_i > cache_layer_id``
    (up).  Up blocks / layers are indexed in reverse, the block-level wrapper has layer 0;
  * ``cur`` = index of the timestep in ``scheduler.timesteps`` (first match), ``start`` = the
    first ``cur`` seen.
PARITY UNPINNED (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class DeepCacheOracle:
    def __init__(self, unet):
        self.unet = unet
        self.set_params()

    def set_params(self, cache_interval=1, cache_branch_id=0, skip_mode="uniform"):
        self.interval = cache_interval
        self.layer_id = cache_branch_id % 3
        self.block_id = cache_branch_id // 3
        self.skip_mode = skip_mode
        self.reset()

    def reset(self):
        self.cache = {}
        self.start = None
        self.cur = 0

    def is_skip_step(self, block_i, layer_i, blocktype="down"):
        if self.start is None:
            self.start = self.cur
        if (self.cur - self.start) % self.interval == 0:
            return False
        if block_i > self.block_id or blocktype == "mid":
            return True
        if block_i < self.block_id:
            return False
        return layer_i >= self.layer_id if blocktype == "down" else layer_i > self.layer_id

    def _wrap(self, key, block_i, layer_i, blocktype, fn):
        if self.is_skip_step(block_i, layer_i, blocktype):
            return self.cache[key]
        out = fn()
        self.cache[key] = out
        return out

    def forward(self, sample, timestep, ctx, cur_index):
        """One UNet call; ``cur_index`` = list(scheduler.timesteps).index(t)."""
        net = self.unet
        self.cur = cur_index
        temb = net.time_embed(sample, timestep)
        h = net.conv_in(sample)
        skips = [h]
        for bi, blk in enumerate(net.down_blocks):
            def run_block(blk=blk, bi=bi, h_in=h):
                hh, outs = h_in, []
                for j in range(len(blk.resnets)):
                    hh = self._wrap(("down", "resnet", bi, j), bi, j, "down",
                                    lambda hh=hh, j=j: blk.resnets[j](hh, temb))
                    if blk.has_attn:
                        hh = self._wrap(("down", "attentions", bi, j), bi, j, "down",
                                        lambda hh=hh, j=j: blk.attentions[j](hh, ctx))
                    outs.append(hh)
                if blk.downsamplers is not None:
                    nres = len(blk.resnets)
                    hh = self._wrap(("down", "downsampler", bi, nres), bi, nres, "down",
                                    lambda hh=hh: blk.downsamplers[0](hh))
                    outs.append(hh)
                return hh, outs

            h, outs = self._wrap(("down", "block", bi, 0), bi, 0, "down", run_block)
            skips += outs

        def run_mid(h_in=h):
            m = net.mid_block
            return m.resnets[1](m.attentions[0](m.resnets[0](h_in, temb), ctx), temb)

        h = self._wrap(("mid", "mid_block", 0, 0), 0, 0, "mid", run_mid)
        nb = len(net.up_blocks)
        for ui, blk in enumerate(net.up_blocks):
            bi = nb - ui - 1
            nl = len(blk.resnets)
            res = skips[-nl:]
            skips = skips[:-nl]

            def run_block(blk=blk, bi=bi, nl=nl, h_in=h, res=res):
                hh, res = h_in, list(res)
                for j in range(nl):
                    li = nl - j - 1
                    skip = res.pop()
                    hh = self._wrap(("up", "resnet", bi, li), bi, li, "up",
                                    lambda hh=hh, j=j, skip=skip: blk.resnets[j](torch.cat([hh, skip], 1), temb))
                    if blk.has_attn:
                        hh = self._wrap(("up", "attentions", bi, li), bi, li, "up",
                                        lambda hh=hh, j=j: blk.attentions[j](hh, ctx))
                if blk.upsamplers is not None:
                    hh = self._wrap(("up", "upsampler", bi, 0), bi, 0, "up", lambda hh=hh: blk.upsamplers[0](hh))
                return hh

            h = self._wrap(("up", "block", bi, 0), bi, 0, "up", run_block)
        return net.conv_out(F.silu(net.conv_norm_out(h)))


def deepcache_forward(net, sample, timestep, ctx, state, full):
    """Convenience for two-call checks: ``full`` step then a branch-0 cached step."""
    helper = state.setdefault("helper", None)
    if helper is None:
        helper = DeepCacheOracle(net)
        helper.set_params(cache_interval=1000, cache_branch_id=0)
        state["helper"] = helper
        state["i"] = 0
    out = helper.forward(sample, timestep, ctx, state["i"] if not full else 0)
    state["i"] += 1
    return out
