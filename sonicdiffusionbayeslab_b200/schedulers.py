"""Scheduler plugins of the B200 engine (registry names of /root/reference/src/schedulers.py).

Registered names (same as the reference, /root/reference/src/schedulers.py:12,190,195):
``dpm_solver_scheduler``, ``ddim_scheduler``, ``lcm_scheduler``; plus ``pndm_scheduler`` for the
stock PNDM/PLMS scheduler the ``default`` / ``deep_cache`` methods run with
(/root/reference/src/experiments/deep_cache.py:17-18).

Split of work
  * host (numpy / 0-dim float32 torch scalars, exactly the arithmetic diffusers uses): timestep
    grids, alphas_cumprod, sigmas, step indices, order / lower-order-final decisions -- these
    are integer / float32 *schedules* and are bit-exact with the reference;
  * device: every ``step`` is ONE launch of the fused latent-update kernel
    (``sonic_latent_update``).  Each scheduler reduces its update rule to the kernel's
    linear-combination coefficients (see include/sonic.h); ``step_cfg`` additionally fuses the
    classifier-free-guidance combine of /root/reference/src/models.py:238-242 into that launch.

The public surface follows diffusers' ``SchedulerMixin`` as used by the reference pipelines
(/root/reference/src/models.py:167-169,222-224,253-255): ``from_config``, ``set_timesteps``,
``timesteps``, ``scale_model_input``, ``step(..., return_dict=False)`` -> tuple, ``order``,
``init_noise_sigma``, ``config``; DPM also exposes ``convert_model_output``, ``model_outputs``,
``step_index``, ``sigmas``, ``lower_order_nums``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from .registry import schedulers_registry


class SchedulerConfig(dict):
    """Mapping with attribute access (diffusers' FrozenDict behaviour)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _betas(cfg) -> torch.Tensor:
    T = cfg["num_train_timesteps"]
    if cfg.get("trained_betas") is not None:
        return torch.tensor(cfg["trained_betas"], dtype=torch.float32)
    sched = cfg["beta_schedule"]
    if sched == "linear":
        return torch.linspace(cfg["beta_start"], cfg["beta_end"], T, dtype=torch.float32)
    if sched == "scaled_linear":
        return torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, T, dtype=torch.float32) ** 2
    raise NotImplementedError(f"{sched} is not implemented")


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """Same draw as diffusers' ``randn_tensor`` (reference: src/schedulers.py:7,139-144)."""
    rand_device = device
    if generator is not None and generator.device.type == "cpu" and torch.device(device).type != "cpu":
        rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)


def _explicit_signature(cls):
    """Expose ``_defaults`` as the keyword signature of ``__init__`` so the registry's
    init-signature dataclass (class_registry.py) lists the real config fields."""
    import inspect

    params = [inspect.Parameter("self", inspect.Parameter.POSITIONAL_OR_KEYWORD)]
    params += [inspect.Parameter(k, inspect.Parameter.KEYWORD_ONLY, default=v) for k, v in cls._defaults.items()]
    cls.__init__.__signature__ = inspect.Signature(params)
    return cls


def _f(x) -> float:
    return float(x.item() if isinstance(x, torch.Tensor) else x)


class FusedScheduler:
    """Common machinery: config handling, step-index bookkeeping, the fused device step."""

    order = 1
    init_noise_sigma = 1.0
    _defaults: dict = {}
    returns_x0 = True

    def __init__(self, **kwargs):
        cfg = dict(self._defaults)
        for k, v in kwargs.items():
            if k not in cfg:
                raise TypeError(f"{type(self).__name__}: unexpected config key {k!r}")
            cfg[k] = v
        self.config = SchedulerConfig(cfg)
        self.betas = _betas(cfg)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)        # float32, CPU
        self.num_inference_steps = None
        self.timesteps = None
        self._timesteps_host = None
        self._step_index = None
        self._replay_device = None          # set by replay(): steps draw their noise and advance, no launches
        self.rng_rows = None                # (row_start, row_stop, global_rows): draw the global tensor, keep rows
        self.x0_rows = None                 # pipelines: write x0 for the leading images only (models.py:257-261)
        self.skip_x0 = False                # pipelines with output_type="latent": x0 predictions are never decoded
        self._ring, self._ring_pos = {}, 0  # history buffers (m0) are recycled, not allocated per step

    # -- construction as in base_experiment.py:69-72: unknown keys are dropped silently
    @classmethod
    def from_config(cls, config, **overrides):
        """diffusers ``ConfigMixin.from_config``: keys this class does not take are not passed to ``__init__`` but
        stay in ``.config`` as hidden entries, so the scheduler-swap idiom of the reference
        (``Other.from_config(pipe.scheduler.config)``, base_experiment.py:69-72) still sees e.g. ``clip_sample=False``
        of the SD-v1.5 scheduler_config.json after a trip through a scheduler that has no such field."""
        merged = dict(config)
        merged.update(overrides)
        obj = cls(**{k: v for k, v in merged.items() if k in cls._defaults})
        for k, v in merged.items():
            if k not in cls._defaults:
                obj.config[k] = v
        return obj

    @property
    def step_index(self):
        return self._step_index

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _set_grid(self, ts: np.ndarray, device):
        self._timesteps_host = [int(t) for t in ts]
        self.timesteps = torch.from_numpy(np.asarray(ts, dtype=np.int64)).to(device)
        self.num_inference_steps = len(ts)
        self._step_index = None

    def _index_for_timestep(self, timestep) -> int:
        t = int(timestep)
        hits = [i for i, v in enumerate(self._timesteps_host) if v == t]
        if not hits:
            return len(self._timesteps_host) - 1
        return hits[1] if len(hits) > 1 else hits[0]

    # -- RNG parity: ONE generator is consumed, in order, by every batch
    # and every noisy step (base_experiment.py:51-53,149; schedulers.py:134-147).
    def _draw(self, sample, generator, dtype):
        """The per-step noise draw of this scheduler, shaped like ``sample``.  Under ``rng_rows`` the GLOBAL
        batch is drawn and this shard's rows are kept (Philox output depends on the tensor shape, so a
        shard-shaped draw would not reproduce the single-process stream)."""
        device = self._replay_device if self._replay_device is not None else sample.device
        shape = tuple(sample.shape)
        if self.rng_rows is None:
            return randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        lo, hi, total = self.rng_rows
        return randn_tensor((total,) + shape[1:], generator=generator, device=device, dtype=dtype)[lo:hi]

    def replay(self, timestep, shape, dtype, device, **step_kwargs):
        """Advance exactly like ``step`` -- same step-index bookkeeping, same draws from ``generator`` -- without
        touching the GPU: how a rank keeps the shared generator aligned over batches another rank computes."""
        meta = torch.empty(shape, dtype=dtype, device="meta")
        self._replay_device = torch.device(device)
        try:
            self._step(meta, None, 0.0, timestep, meta, **step_kwargs)
        finally:
            self._replay_device = None

    # -- the device step
    def _launch(self, coeffs, eps, eps_text, sample, hist=(), noise=None, want_m0=False, want_x0=True, out=None,
                ring=True, post=None):
        """One fused launch.  ``post``: non-linear x0 post-processing (``dict(mode=1, clip=)`` for ``clip_sample``,
        ``dict(mode=2, ratio=, max_value=)`` for dynamic thresholding, plus ``p_x`` / ``p_0``) -- the thresholding
        scale is one extra launch (a per-image quantile) in front of the update."""
        if self._replay_device is not None:
            return sample, (sample if want_m0 else None), (sample if want_x0 else None)
        if not sample.is_cuda:
            raise RuntimeError(f"{type(self).__name__}.step needs CUDA tensors: the B200 engine has no CPU path")
        sample = sample.contiguous()
        out_sample = torch.empty_like(sample) if out is None else out
        out_m0 = (self._history_buffer(sample) if ring else torch.empty_like(sample)) if want_m0 else None
        out_x0 = None
        if want_x0 and not self.skip_x0:
            rows = sample.shape[0] if self.x0_rows is None else min(self.x0_rows, sample.shape[0])
            out_x0 = torch.empty((rows,) + tuple(sample.shape[1:]), dtype=sample.dtype, device=sample.device)
        h = list(hist) + [None] * (3 - len(hist))
        eps = eps.contiguous()
        if post is not None and post["mode"] == 2:
            post = dict(post, thr=K.x0_threshold(coeffs, eps, sample, eps_text=eps_text, ratio=post["ratio"],
                                                 max_value=post["max_value"]))
        K.latent_update(coeffs, eps, sample, eps_text=eps_text, h1=h[0], h2=h[1], h3=h[2], noise=noise,
                        out_sample=out_sample, out_m0=out_m0, out_x0=out_x0, post=post)
        return out_sample, out_m0, out_x0

    def _x0_post(self, p_x=0.0, p_0=1.0):
        """The config's x0 post-processing as the ``post`` argument of ``_launch`` (None: the linear hot path).
        Thresholding wins over clipping, as in diffusers (``if thresholding ... elif clip_sample``)."""
        cfg = self.config
        if cfg.get("thresholding"):
            return dict(mode=2, ratio=cfg.get("dynamic_thresholding_ratio", 0.995),
                        max_value=cfg.get("sample_max_value", 1.0), p_x=p_x, p_0=p_0)
        if cfg.get("clip_sample"):
            return dict(mode=1, clip=cfg.get("clip_sample_range", 1.0), p_x=p_x, p_0=p_0)
        return None

    _RING = 6            # > the longest history any scheduler keeps alive (PNDM: 4 ets + the one being written)

    def _history_buffer(self, like):
        """Next slot of a small ring of latent-shaped buffers: the converted model outputs kept as multistep
        history live for at most ``solver_order`` (DPM) / 4 (PLMS) steps, so nothing is allocated per step."""
        key = (tuple(like.shape), like.dtype, like.device)
        ring = self._ring.get(key)
        if ring is None:
            ring = self._ring[key] = [torch.empty_like(like) for _ in range(self._RING)]
        self._ring_pos = (self._ring_pos + 1) % self._RING
        return ring[self._ring_pos]

    def step(self, model_output, timestep, sample, return_dict=False, **kw):
        return self._step(model_output, None, 0.0, timestep, sample, **kw)

    def step_cfg(self, eps_uncond, eps_text, guidance_scale, timestep, sample, **kw):
        """Fused ``uncond + g*(text-uncond)`` + ``step`` (src/models.py:238-242 + 253-255)."""
        return self._step(eps_uncond, eps_text, guidance_scale, timestep, sample, **kw)


# ======================================================================================= DDIM
@schedulers_registry.add_to_registry("ddim_scheduler")
@_explicit_signature
class DDIMSchedulerMy(FusedScheduler):
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", trained_betas=None,
        clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", thresholding=False,
        dynamic_thresholding_ratio=0.995, clip_sample_range=1.0, sample_max_value=1.0, timestep_spacing="leading",
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.config.prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, "
                             "`sample`, or `v_prediction`")
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self._set_grid(np.arange(0, self.config.num_train_timesteps)[::-1].copy(), None)
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps, device=None, **_):
        T = self.config.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {T}")
        sp = self.config.timestep_spacing
        if sp == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif sp == "trailing":
            ts = np.round(np.arange(T, 0, -T / num_inference_steps)).astype(np.int64) - 1
        elif sp == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        else:
            raise ValueError(f"{sp} is not supported")
        self._set_grid(ts, device)

    def _step(self, eps, eps_text, guidance, timestep, sample, eta: float = 0.0, generator=None, out=None, **_):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta = 1 - a
        var = ((1 - ap) / beta) * (1 - a / ap)
        std = eta * var ** 0.5
        sa, sb, direction = _f(a ** 0.5), _f(beta ** 0.5), _f((1 - ap - std ** 2) ** 0.5)
        # prev = sqrt(ap) x0 + direction * pred_eps, with x0 = m (rounded to the model dtype like the reference's
        # pred_original_sample tensor) and pred_eps linear in (model output e, x) for every prediction type
        pt = self.config.prediction_type
        if pt == "epsilon":
            c = dict(m_x=1.0 / sa, m_e=-sb * (1.0 / sa), c_e=direction)
        elif pt == "sample":
            c = dict(m_x=0.0, m_e=1.0, c_x=direction / sb, c_e=-direction * sa / sb)
        else:                                             # v_prediction
            c = dict(m_x=sa, m_e=-sb, c_x=direction * sb, c_e=direction * sa)
        c.update(guidance=guidance, c_m0=_f(ap ** 0.5), x0_x=c["m_x"], x0_e=c["m_e"])
        noise = None
        if eta > 0:
            noise = self._draw(sample, generator, sample.dtype)
            c["c_z"] = _f(std)
        # clip_sample / thresholding act on pred_original_sample only (use_clipped_model_output is False): m0 = x0'
        prev, _, x0 = self._launch(c, eps, eps_text, sample, noise=noise, out=out, post=self._x0_post())
        return (prev, x0)


# ======================================================================================= DPM-Solver
@schedulers_registry.add_to_registry("dpm_solver_scheduler")
@_explicit_signature
class DPMSolverScheduler(FusedScheduler):
    """Multistep DPM-Solver(++) with the semantics of /root/reference/src/schedulers.py:12-187:
    ``step`` returns ``(prev_sample, x0_pred)``; ``convert_model_output`` returns
    ``(epsilon, x0_pred)`` for the non-``++`` algorithms (:96) and, for ``++`` -- where the
    reference code cannot run -- the intended ``(x0_pred, x0_pred)``."""

    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", trained_betas=None,
        solver_order=2, prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995,
        sample_max_value=1.0, algorithm_type="dpmsolver++", solver_type="midpoint", lower_order_final=True,
        euler_at_final=False, use_karras_sigmas=False,
        lambda_min_clipped=-float("inf"), variance_type=None, timestep_spacing="linspace", steps_offset=0,
        final_sigmas_type="zero",
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        cfg = self.config
        if cfg.algorithm_type not in ("dpmsolver", "dpmsolver++", "sde-dpmsolver", "sde-dpmsolver++"):
            if cfg.algorithm_type == "deis":
                cfg["algorithm_type"] = "dpmsolver++"
            else:
                raise NotImplementedError(f"{cfg.algorithm_type} is not implemented for {self.__class__}")
        if cfg.solver_type not in ("midpoint", "heun"):
            if cfg.solver_type in ("logrho", "bh1", "bh2"):
                cfg["solver_type"] = "midpoint"
            else:
                raise NotImplementedError(f"{cfg.solver_type} is not implemented for {self.__class__}")
        if cfg.algorithm_type not in ("dpmsolver++", "sde-dpmsolver++") and cfg.final_sigmas_type == "zero":
            raise ValueError(f"`final_sigmas_type` {cfg.final_sigmas_type} is not supported for "
                             f"`algorithm_type` {cfg.algorithm_type}. Please choose `sigma_min` instead.")
        pp = cfg.algorithm_type in ("dpmsolver++", "sde-dpmsolver++")
        allowed = ("epsilon", "sample", "v_prediction") + (("flow_prediction",) if pp else ())
        if cfg.prediction_type not in allowed:                   # src/schedulers.py:52-56 / :80-83
            raise ValueError(f"prediction_type given as {cfg.prediction_type} must be one of `epsilon`, `sample`, "
                             + ("`v_prediction`, or `flow_prediction`" if pp else "or `v_prediction`")
                             + " for the DPMSolverMultistepScheduler.")
        if cfg.use_karras_sigmas:
            raise NotImplementedError("fused DPM-Solver step: no Karras sigmas (not used by any reference config)")
        if not 1 <= cfg.solver_order <= 3:
            raise NotImplementedError("solver_order must be 1, 2 or 3")
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.model_outputs = [None] * cfg.solver_order
        self.lower_order_nums = 0

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if num_inference_steps is None and timesteps is None:
            raise ValueError("Must pass exactly one of `num_inference_steps` or `timesteps`.")
        cfg = self.config
        T = cfg.num_train_timesteps
        if timesteps is not None:
            ts = np.array(timesteps).astype(np.int64)
        else:
            clipped = torch.searchsorted(torch.flip(self.lambda_t, [0]), cfg.lambda_min_clipped)
            last = int((T - clipped).item())
            if cfg.timestep_spacing == "linspace":
                ts = np.linspace(0, last - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
            elif cfg.timestep_spacing == "leading":
                ratio = last // (num_inference_steps + 1)
                ts = (np.arange(0, num_inference_steps + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64)
                ts += cfg.steps_offset
            elif cfg.timestep_spacing == "trailing":
                ts = np.arange(last, 0, -T / num_inference_steps).round().copy().astype(np.int64) - 1
            else:
                raise ValueError(f"{cfg.timestep_spacing} is not supported")
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        if cfg.final_sigmas_type == "sigma_min":
            sigma_last = float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5)
        elif cfg.final_sigmas_type == "zero":
            sigma_last = 0
        else:
            raise ValueError(f"`final_sigmas_type` must be one of 'zero', or 'sigma_min', but got "
                             f"{cfg.final_sigmas_type}")
        self.sigmas = torch.from_numpy(np.concatenate([sig, [sigma_last]]).astype(np.float32))
        self._set_grid(ts, device)
        self.model_outputs = [None] * cfg.solver_order
        self.lower_order_nums = 0

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _asl(self, sigma):
        a, s = self._sigma_to_alpha_sigma_t(sigma)
        return a, s, torch.log(a) - torch.log(s)

    def _convert_coeffs(self):
        """``convert_model_output`` (src/schedulers.py:14-96) as the fused kernel's two linear forms of (x, model
        output e): the converted output kept as multistep history, ``m = m_x x + m_e e``, and the returned x0
        prediction ``x0 = x0_x x + x0_e e`` -- every ``prediction_type`` branch is linear, so none needs a kernel of
        its own.  Scalars are float32 exactly as the reference computes them (sigma -> alpha_t, sigma_t), combined in
        Python floats; the non-``++`` x0 is ``(x - sigma_t eps) / alpha_t`` (:92-94) with eps substituted."""
        sigma = self.sigmas[self.step_index]
        a_s, s_s = self._sigma_to_alpha_sigma_t(sigma)
        a, s = _f(a_s), _f(s_s)
        pt = self.config.prediction_type
        if self.config.algorithm_type in ("dpmsolver++", "sde-dpmsolver++"):
            if pt == "epsilon":
                x0_x, x0_e = 1.0 / a, -s / a                     # :40-42
            elif pt == "sample":
                x0_x, x0_e = 0.0, 1.0                            # :43-44
            elif pt == "v_prediction":
                x0_x, x0_e = a, -s                               # :45-48
            else:                                                # flow_prediction, :49-51: sigma itself
                x0_x, x0_e = 1.0, -_f(sigma)
            return dict(m_x=x0_x, m_e=x0_e, x0_x=x0_x, x0_e=x0_e)
        if pt == "epsilon":
            m_x, m_e = 0.0, 1.0                                  # :66-71
        elif pt == "sample":
            m_x, m_e = 1.0 / s, -a / s                           # :72-75
        else:                                                    # v_prediction, :76-79
            m_x, m_e = s, a
        return dict(m_x=m_x, m_e=m_e, x0_x=(1.0 - s * m_x) / a, x0_e=-s * m_e / a)

    def _threshold_post(self):
        """``thresholding`` (src/schedulers.py:58-59 / :85-90): x0 is dynamically thresholded and the converted model
        output re-derived from it -- x0' itself for the ``++`` algorithms, ``(x - alpha_t x0') / sigma_t`` otherwise."""
        if not self.config.thresholding:
            return None
        if self.config.algorithm_type in ("dpmsolver++", "sde-dpmsolver++"):
            return self._x0_post(0.0, 1.0)
        a_s, s_s = self._sigma_to_alpha_sigma_t(self.sigmas[self.step_index])
        return self._x0_post(1.0 / _f(s_s), -_f(a_s) / _f(s_s))

    def convert_model_output(self, model_output, *args, sample=None, **kwargs):
        """src/schedulers.py:14-96 -> (converted model output, x0_pred); one fused launch."""
        if sample is None:
            if len(args) > 1:
                sample = args[1]
            else:
                raise ValueError("missing `sample` as a required keyward argument")
        if self.step_index is None:
            raise ValueError("convert_model_output needs an initialised step index (call step first)")
        c = self._convert_coeffs()
        c["c_x"] = 1.0                                            # out_sample is a scratch copy of x
        _, m0, x0 = self._launch(c, model_output, None, sample, want_m0=True, ring=False,   # caller owns the result
                                 post=self._threshold_post())
        return m0, x0

    def feed_history(self, eps, eps_text, guidance, sample):
        """``model_outputs <- shift + convert_model_output(uncond + g (text - uncond), sample=sample)`` at the CURRENT
        step index, without stepping: how the interleaved pipeline keeps this scheduler's multistep history alive
        while the other scheduler advances the latents (src/models.py:1024-1031, 1045-1053).  One fused launch."""
        if self.step_index is None:
            raise ValueError("feed_history needs an initialised step index (src/schedulers.py:40 indexes "
                             "sigmas[None] in the reference: the main scheduler has to take a step first)")
        c = dict(guidance=guidance, **self._convert_coeffs())
        c["c_x"] = 1.0
        _, m0, _ = self._launch(c, eps, eps_text, sample, want_m0=True, want_x0=False, post=self._threshold_post())
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m0

    def _update_coeffs(self, order):
        """Linear-combination form of dpm_solver_first_order_update /
        multistep_dpm_solver_{second,third}_order_update (diffusers 0.32.1)."""
        cfg, i = self.config, self.step_index
        algo, heun = cfg.algorithm_type, cfg.solver_type == "heun"
        a_t, s_t, l_t = self._asl(self.sigmas[i + 1])
        a_s0, s_s0, l_s0 = self._asl(self.sigmas[i])
        h = l_t - l_s0
        pp = algo in ("dpmsolver++", "sde-dpmsolver++")
        sde = algo.startswith("sde")
        c = {}
        if not sde:
            if pp:
                c["c_x"] = _f(s_t / s_s0)
                A = _f(a_t * (torch.exp(-h) - 1.0))               # x' = c_x x - A D0 ...
            else:
                c["c_x"] = _f(a_t / a_s0)
                A = _f(s_t * (torch.exp(h) - 1.0))
            d0, d1, d2 = -A, 0.0, 0.0
            if order >= 2:
                if order == 2 and not heun:
                    d1 = -0.5 * A
                elif pp:
                    d1 = _f(a_t * ((torch.exp(-h) - 1.0) / h + 1.0))
                else:
                    d1 = -_f(s_t * ((torch.exp(h) - 1.0) / h - 1.0))
            if order == 3:
                if pp:
                    d2 = -_f(a_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5))
                else:
                    d2 = -_f(s_t * ((torch.exp(h) - 1.0 - h) / h ** 2 - 0.5))
        else:
            if order == 3:
                raise NotImplementedError("third-order SDE DPM-Solver does not exist in diffusers")
            if pp:
                c["c_x"] = _f(s_t / s_s0 * torch.exp(-h))
                Bc = _f(a_t * (1 - torch.exp(-2.0 * h)))
                d0 = Bc
                d1 = 0.0 if order == 1 else (0.5 * Bc if not heun else
                                             _f(a_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)))
                c["c_z"] = _f(s_t * torch.sqrt(1.0 - torch.exp(-2.0 * h)))
            else:
                c["c_x"] = _f(a_t / a_s0)
                S = _f(s_t * (torch.exp(h) - 1.0))
                d0 = -2.0 * S
                d1 = 0.0 if order == 1 else (-S if not heun else
                                             -2.0 * _f(s_t * ((torch.exp(h) - 1.0) / h - 1.0)))
                c["c_z"] = _f(s_t * torch.sqrt(torch.exp(2.0 * h) - 1.0))
            d2 = 0.0
        # D0 = m0 ; D1_0 = (m0-m1)/r0 ; D1_1 = (m1-m2)/r1 ; order 2: D1 = D1_0
        # order 3: D1 = D1_0 + r0/(r0+r1) (D1_0-D1_1) ; D2 = (D1_0-D1_1)/(r0+r1)
        c_m0, c_h1, c_h2 = d0, 0.0, 0.0
        if order >= 2:
            _, _, l_s1 = self._asl(self.sigmas[i - 1])
            r0 = _f((l_s0 - l_s1) / h)
            if order == 2:
                P, Q = d1, 0.0
                r1 = 1.0
            else:
                _, _, l_s2 = self._asl(self.sigmas[i - 2])
                r1 = _f((l_s1 - l_s2) / h)
                rho = r0 / (r0 + r1)
                P = d1 * (1.0 + rho) + d2 / (r0 + r1)
                Q = -d1 * rho - d2 / (r0 + r1)
            c_m0 += P / r0
            c_h1 += -P / r0 + Q / r1
            c_h2 += -Q / r1
        c.update(c_m0=c_m0, c_h1=c_h1, c_h2=c_h2)
        return c

    def _step(self, eps, eps_text, guidance, timestep, sample, generator=None, variance_noise=None, out=None, **_):
        """src/schedulers.py:98-187 in one launch."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        if self.step_index is None:
            self._step_index = self._index_for_timestep(timestep)
        cfg = self.config
        n = len(self._timesteps_host)
        lower_order_final = (self.step_index == n - 1) and (
            cfg.euler_at_final or (cfg.lower_order_final and n < 15) or cfg.final_sigmas_type == "zero")
        lower_order_second = (self.step_index == n - 2) and cfg.lower_order_final and n < 15
        if cfg.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            order = 1
        elif cfg.solver_order == 2 or self.lower_order_nums < 2 or lower_order_second:
            order = 2
        else:
            order = 3
        c = dict(guidance=guidance, **self._convert_coeffs())
        c.update(self._update_coeffs(order))
        noise = None
        if cfg.algorithm_type in ("sde-dpmsolver", "sde-dpmsolver++"):
            if variance_noise is None:
                noise = self._draw(sample, generator, torch.float32)
            else:
                noise = variance_noise.to(device=sample.device, dtype=torch.float32)
            noise = noise.to(sample.dtype)
        hist = [m for m in (self.model_outputs[-1], self.model_outputs[-2] if cfg.solver_order > 1 else None)
                if m is not None][: order - 1]
        # history as seen by this step: model_outputs[-1] is m1 (previous), [-2] is m2
        prev, m0, x0 = self._launch(c, eps, eps_text, sample, hist=hist, noise=noise, want_m0=True, out=out,
                                    post=self._threshold_post())
        for i in range(cfg.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m0
        if self.lower_order_nums < cfg.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        return (prev, x0)


# ======================================================================================= LCM
@schedulers_registry.add_to_registry("lcm_scheduler")
@_explicit_signature
class LCMScheduler(FusedScheduler):
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
        trained_betas=None, original_inference_steps=50, clip_sample=False, clip_sample_range=1.0,
        set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", thresholding=False,
        timestep_spacing="leading", timestep_scaling=10.0,
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.config.prediction_type not in ("epsilon", "sample", "v_prediction"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, "
                             "`sample` or `v_prediction`")
        if self.config.thresholding or self.config.clip_sample:
            raise NotImplementedError("fused LCM step: no clipping / thresholding of x0 (off in the LCM configuration "
                                      "of the reference)")
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.sigma_data = 0.5

    def set_timesteps(self, num_inference_steps, device=None, original_inference_steps=None, **_):
        T = self.config.num_train_timesteps
        orig = original_inference_steps or self.config.original_inference_steps
        if orig > T:
            raise ValueError(f"`original_steps`: {orig} cannot be larger than {T}")
        if num_inference_steps > orig:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`original_inference_steps`: {orig}")
        k = T // orig
        origin = (np.asarray(list(range(1, orig + 1))) * k - 1)[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        self._set_grid(origin[idx], device)

    def get_scalings_for_boundary_condition_discrete(self, timestep):
        scaled = torch.as_tensor(timestep).cpu() * self.config.timestep_scaling
        c_skip = self.sigma_data ** 2 / (scaled ** 2 + self.sigma_data ** 2)
        c_out = scaled / (scaled ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def _step(self, eps, eps_text, guidance, timestep, sample, generator=None, out=None, **_):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        if self.step_index is None:
            self._step_index = self._index_for_timestep(timestep)
        t = int(timestep)
        nxt = self.step_index + 1
        prev_t = self._timesteps_host[nxt] if nxt < len(self._timesteps_host) else t
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(torch.tensor(t))
        sa, sb = _f(a.sqrt()), _f((1 - a).sqrt())
        pt = self.config.prediction_type                             # predicted_original_sample = m_x x + m_e e
        m_x, m_e = (1.0 / sa, -sb * (1.0 / sa)) if pt == "epsilon" else (0.0, 1.0) if pt == "sample" else (sa, -sb)
        c = dict(guidance=guidance, m_x=m_x, m_e=m_e,
                 x0_x=_f(c_out) * m_x + _f(c_skip), x0_e=_f(c_out) * m_e)   # denoised
        last = self.step_index == self.num_inference_steps - 1
        noise = None
        if not last:
            noise = self._draw(sample, generator, sample.dtype)
            s = _f(ap.sqrt())
            c.update(c_m0=s * _f(c_out), c_x=s * _f(c_skip), c_z=_f((1 - ap).sqrt()))
        else:
            c.update(c_m0=_f(c_out), c_x=_f(c_skip))
        prev, _, den = self._launch(c, eps, eps_text, sample, noise=noise, out=out)
        self._step_index += 1
        return (prev, den)


# ======================================================================================= PNDM / PLMS
@schedulers_registry.add_to_registry("pndm_scheduler")
@_explicit_signature
class PNDMScheduler(FusedScheduler):
    returns_x0 = False
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", trained_betas=None,
        skip_prk_steps=False, set_alpha_to_one=False, prediction_type="epsilon", timestep_spacing="leading",
        steps_offset=0,
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        if self.config.prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon` or "
                             "`v_prediction`")
        if not self.config.skip_prk_steps:
            raise NotImplementedError("only the PLMS path (skip_prk_steps=True) of SD-v1.5 is fused")
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.ets, self.counter, self.cur_sample = [], 0, None

    def set_timesteps(self, num_inference_steps, device=None, **_):
        T = self.config.num_train_timesteps
        if self.config.timestep_spacing != "leading":
            raise NotImplementedError("PNDM: only leading spacing (SD-v1.5) is implemented")
        ratio = T // num_inference_steps
        _t = (np.arange(0, num_inference_steps) * ratio).round()
        _t += self.config.steps_offset
        plms = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy()
        self._set_grid(plms.astype(np.int64), device)
        self.num_inference_steps = num_inference_steps
        self.ets, self.counter, self.cur_sample = [], 0, None

    def _step(self, eps, eps_text, guidance, timestep, sample, out=None, **_):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        t = int(timestep)
        stride = self.config.num_train_timesteps // self.num_inference_steps
        prev_t = t - stride
        append = self.counter != 1
        if not append:
            prev_t, t = t, t + stride
        n_hist = len(self.ets[-3:]) if append else len(self.ets)
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta, beta_prev = 1 - a, 1 - ap
        sample_coeff = (ap / a) ** 0.5
        denom = a * beta_prev ** 0.5 + (a * beta * ap) ** 0.5
        kappa = _f((ap - a) / denom)
        c = dict(guidance=guidance, m_x=0.0, m_e=1.0, c_x=_f(sample_coeff))
        hist = []
        if append and n_hist == 0:                                   # first call
            c["c_m0"] = -kappa
        elif not append:                                             # repeated timestep: average, stored sample
            c["c_m0"], c["c_h1"] = -0.5 * kappa, -0.5 * kappa
            hist = [self.ets[-1]]
        elif n_hist == 1:
            c["c_m0"], c["c_h1"] = -1.5 * kappa, 0.5 * kappa
            hist = [self.ets[-1]]
        elif n_hist == 2:
            c["c_m0"], c["c_h1"], c["c_h2"] = -23 / 12 * kappa, 16 / 12 * kappa, -5 / 12 * kappa
            hist = [self.ets[-1], self.ets[-2]]
        else:
            c["c_m0"], c["c_h1"], c["c_h2"], c["c_h3"] = (-55 / 24 * kappa, 59 / 24 * kappa, -37 / 24 * kappa,
                                                          9 / 24 * kappa)
            hist = [self.ets[-1], self.ets[-2], self.ets[-3]]
        if self.config.prediction_type == "v_prediction":
            # _get_prev_sample turns the combined v outputs into a noise prediction with the CURRENT sample:
            # eps = sqrt(a) v + sqrt(beta) x  ->  every history weight scales by sqrt(a), x gains -kappa sqrt(beta)
            for key in ("c_m0", "c_h1", "c_h2", "c_h3"):
                if key in c:
                    c[key] *= _f(a ** 0.5)
            c["c_x"] -= kappa * _f(beta ** 0.5)
        src = sample
        if not append:
            src = self.cur_sample
            self.cur_sample = None
        elif n_hist == 0:
            # the stored sample must survive an in-place update of `sample` (out aliasing it)
            self.cur_sample = sample.clone() if out is not None and out.data_ptr() == sample.data_ptr() else sample
        prev, m0, _ = self._launch(c, eps, eps_text, src, hist=hist, want_m0=append, want_x0=False, out=out)
        if append:
            self.ets = self.ets[-3:]
            self.ets.append(m0)
        self.counter += 1
        return (prev,)
