"""DeepCache feature reuse for the B200 engine.

Drop-in for ``DeepCache.DeepCacheSDHelper`` as driven by the reference
(/root/reference/src/experiments/deep_cache.py:24-29,58): ``DeepCacheSDHelper(pipe=model)``,
``set_params(cache_interval=, cache_branch_id=)``, ``enable()``, ``disable()``.

Instead of wrapping ~90 module forwards with python dict lookups, enabling it makes the pipeline
replay the engine's *cached* launch plan on every step that is not a refresh step:
step ``i`` is a full step iff ``(cur - start) % cache_interval == 0`` with
``cur = list(timesteps).index(t)``; a cached step recomputes only time-MLP, conv_in, the last up
resnet + transformer, norm_out and conv_out (63.25 GFLOP/sample instead of 803.27) and reads the
skip-branch feature -- the output of ``up_blocks[-1].attentions[1]`` -- that the last full step
left resident in HBM.  Other ``cache_branch_id`` values move the cut: ``block_id, layer_id =
divmod(branch, 3)`` selects how many down layers are recomputed and which up layer takes the cached
feature as its input (``UNetEngine._build_unet_plan``); the engine records one cached plan per branch.
"""
from __future__ import annotations


class DeepCacheSDHelper:
    def __init__(self, pipe=None):
        self.pipe = pipe
        self.params = None

    def set_params(self, cache_interval=1, cache_branch_id=0, skip_mode="uniform"):
        if skip_mode != "uniform":
            raise NotImplementedError("only skip_mode='uniform' (the reference's setting) is implemented")
        if not 0 <= int(cache_branch_id) < 12:
            raise ValueError("cache_branch_id must be in [0, 12): 4 blocks x 3 layers (DeepCacheSDHelper.set_params)")
        if cache_interval < 1:
            raise ValueError("cache_interval must be >= 1")
        self.params = dict(interval=int(cache_interval), branch=int(cache_branch_id), start=None)
        return self

    def enable(self, pipe=None):
        if pipe is not None:
            self.pipe = pipe
        if self.pipe is None or self.params is None:
            raise RuntimeError("call DeepCacheSDHelper(pipe=...) and set_params(...) before enable()")
        self.pipe._deepcache = dict(self.params)

    def disable(self):
        if self.pipe is not None:
            self.pipe._deepcache = None
