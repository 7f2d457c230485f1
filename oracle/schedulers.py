"""Oracle schedulers: literal PyTorch restatement (test infrastructure only).

Restates, op by op, the scheduler arithmetic the reference reaches through
diffusers 0.32.1 (absent from /root/reference; pinned at poetry.lock:454-455):

  * ``DDIMScheduler``   <- /root/reference/src/schedulers.py:190-192 (``DDIMSchedulerMy``)
  * ``DPMSolverScheduler`` <- /root/reference/src/schedulers.py:12-187 (the repo's own
    ``convert_model_output`` / ``step`` override on top of DPMSolverMultistepScheduler)
  * ``LCMScheduler``    <- /root/reference/src/schedulers.py:195-197
  * ``PNDMScheduler``   <- stock scheduler used by deep_cache.py:17-18 / default_sd.py:15-16

Every tensor op is kept as a separate ATen op in the same order and dtype as the
published algorithm, so per-op rounding (model dtype for DDIM/LCM/PNDM, the fp32
upcast of schedulers.py:133 for DPM) is reproduced.  Scalars are 0-dim float32 CPU
tensors exactly as diffusers holds them.  Pinned against the reference's own source where it has one (oracle/refexec.py; see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

SD15_SCHEDULER_CONFIG = dict(
    # runwayml/stable-diffusion-v1-5 scheduler/scheduler_config.json (PNDM) + instantiated defaults
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    trained_betas=None,
    set_alpha_to_one=False,
    skip_prk_steps=True,
    steps_offset=1,
    clip_sample=False,
    prediction_type="epsilon",
    timestep_spacing="leading",
)


class _Config(dict):
    __getattr__ = dict.__getitem__


def make_betas(cfg) -> torch.Tensor:
    T = cfg["num_train_timesteps"]
    if cfg.get("trained_betas") is not None:
        return torch.tensor(cfg["trained_betas"], dtype=torch.float32)
    if cfg["beta_schedule"] == "linear":
        return torch.linspace(cfg["beta_start"], cfg["beta_end"], T, dtype=torch.float32)
    if cfg["beta_schedule"] == "scaled_linear":
        return torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, T, dtype=torch.float32) ** 2
    raise NotImplementedError(cfg["beta_schedule"])


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor for a single generator."""
    rand_device = device
    if generator is not None:
        gen_device = generator.device.type
        if gen_device != torch.device(device).type and gen_device == "cpu":
            rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)


class _SchedulerBase:
    order = 1
    init_noise_sigma = 1.0
    _defaults: dict = {}

    def __init__(self, **kwargs):
        cfg = dict(self._defaults)
        unknown = set(kwargs) - set(cfg)
        if unknown:
            raise TypeError(f"unexpected config keys {sorted(unknown)}")
        cfg.update(kwargs)
        self.config = _Config(cfg)
        self.betas = make_betas(cfg)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.num_inference_steps = None
        self.timesteps = None
        self._step_index = None

    @classmethod
    def from_config(cls, config, **overrides):
        """ConfigMixin.from_config: accepted keys go to ``__init__``; the rest stay in ``.config`` as hidden
        entries (so ``Other.from_config(this.config)`` still sees them)."""
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

    def _threshold_sample(self, sample):
        """diffusers 0.32.1 ``_threshold_sample`` ("dynamic thresholding", Imagen): per sample the
        ``dynamic_thresholding_ratio`` quantile s of |x0|, clamped to [1, sample_max_value]; x0 <- clamp(x0, -s, s) / s.
        Third-party restatement; reached from /root/reference/src/schedulers.py:58-59,85-90 (off in every shipped
        config)."""
        dtype = sample.dtype
        b = sample.shape[0]
        flat = sample.float().reshape(b, -1)
        s = torch.quantile(flat.abs(), self.config.get("dynamic_thresholding_ratio", 0.995), dim=1)
        s = torch.clamp(s, min=1, max=self.config.get("sample_max_value", 1.0)).unsqueeze(1)
        flat = torch.clamp(flat, -s, s) / s
        return flat.reshape(sample.shape).to(dtype)

    def index_for_timestep(self, timestep):
        cand = (self.timesteps == timestep).nonzero()
        if len(cand) == 0:
            return len(self.timesteps) - 1
        if len(cand) > 1:
            return cand[1].item()
        return cand[0].item()

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        self._step_index = self.index_for_timestep(timestep)


# ----------------------------------------------------------------------------- DDIM


class DDIMScheduler(_SchedulerBase):
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
        trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
        prediction_type="epsilon", thresholding=False, dynamic_thresholding_ratio=0.995, clip_sample_range=1.0,
        sample_max_value=1.0, timestep_spacing="leading",
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        T = self.config.num_train_timesteps
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None, **_):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        if self.config.timestep_spacing == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
            ts += self.config.steps_offset
        elif self.config.timestep_spacing == "trailing":
            ratio = T / num_inference_steps
            ts = np.round(np.arange(T, 0, -ratio)).astype(np.int64) - 1
        else:  # linspace
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def _get_variance(self, t, prev_t):
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return ((1 - ap) / (1 - a)) * (1 - a / ap)

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, return_dict=False):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta = 1 - a
        pt = self.config.prediction_type                  # diffusers 0.32.1 scheduling_ddim.py ``step`` item 3
        if pt == "epsilon":
            pred_x0 = (sample - beta ** 0.5 * model_output) / a ** 0.5
            pred_eps = model_output
        elif pt == "sample":
            pred_x0 = model_output
            pred_eps = (sample - a ** 0.5 * pred_x0) / beta ** 0.5
        elif pt == "v_prediction":
            pred_x0 = (a ** 0.5) * sample - (beta ** 0.5) * model_output
            pred_eps = (a ** 0.5) * model_output + (beta ** 0.5) * sample
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample`, or `v_prediction`")
        if self.config.thresholding:                       # step item 4: thresholding wins over clipping
            pred_x0 = self._threshold_sample(pred_x0)
        elif self.config.clip_sample:
            pred_x0 = pred_x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
        var = self._get_variance(t, prev_t)
        std = eta * var ** 0.5
        direction = (1 - ap - std ** 2) ** 0.5 * pred_eps
        prev = ap ** 0.5 * pred_x0 + direction
        if eta > 0:
            z = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                             dtype=model_output.dtype)
            prev = prev + std * z
        return (prev, pred_x0)


# ----------------------------------------------------------------------------- DPM-Solver multistep


class DPMSolverScheduler(_SchedulerBase):
    """DPMSolverMultistepScheduler + the override of /root/reference/src/schedulers.py:12-187.

    Quirk C-1: for ``dpmsolver++`` the reference's
    ``convert_model_output`` returns ONE tensor that ``step`` unpacks as two;
    the intended semantics implemented here: solver input = x0_pred, returned
    x0_pred = x0_pred.
    """

    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
        trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
        dynamic_thresholding_ratio=0.995, sample_max_value=1.0, algorithm_type="dpmsolver++", solver_type="midpoint", lower_order_final=True,
        euler_at_final=False, use_karras_sigmas=False, lambda_min_clipped=-float("inf"),
        variance_type=None, timestep_spacing="linspace", steps_offset=0, final_sigmas_type="zero",
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        at = self.config.algorithm_type
        if at not in ("dpmsolver", "dpmsolver++", "sde-dpmsolver", "sde-dpmsolver++"):
            if at == "deis":
                self.config["algorithm_type"] = "dpmsolver++"
            else:
                raise NotImplementedError(f"{at} is not implemented for {self.__class__}")
        if self.config.solver_type not in ("midpoint", "heun"):
            raise NotImplementedError(self.config.solver_type)
        if self.config.algorithm_type not in ("dpmsolver++", "sde-dpmsolver++") and \
                self.config.final_sigmas_type == "zero":
            raise ValueError("`final_sigmas_type` zero is not supported for `algorithm_type` "
                             f"{self.config.algorithm_type}. Please choose `sigma_min` instead.")
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        T = self.config.num_train_timesteps
        self.timesteps = torch.from_numpy(np.linspace(0, T - 1, T, dtype=np.float32)[::-1].copy())

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        T = self.config.num_train_timesteps
        if timesteps is not None:
            ts = np.array(timesteps).astype(np.int64)
        else:
            clipped = torch.searchsorted(torch.flip(self.lambda_t, [0]), self.config.lambda_min_clipped)
            last = int((T - clipped).item())
            sp = self.config.timestep_spacing
            if sp == "linspace":
                ts = np.linspace(0, last - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
            elif sp == "leading":
                ratio = last // (num_inference_steps + 1)
                ts = (np.arange(0, num_inference_steps + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64)
                ts += self.config.steps_offset
            else:
                ratio = T / num_inference_steps
                ts = np.arange(last, 0, -ratio).round().copy().astype(np.int64) - 1
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        if self.config.final_sigmas_type == "sigma_min":
            last_sigma = float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5)
        elif self.config.final_sigmas_type == "zero":
            last_sigma = 0
        else:
            raise ValueError(self.config.final_sigmas_type)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [last_sigma]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self._step_index = None

    @staticmethod
    def _sigma_to_alpha_sigma_t(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def convert_model_output(self, model_output, sample):
        """/root/reference/src/schedulers.py:14-96: every ``prediction_type`` branch (:36-56 for the ``++`` algorithms,
        :65-83 for the others) and the thresholding branches (:58-59, :85-90)."""
        pt = self.config.prediction_type
        sigma = self.sigmas[self.step_index]
        alpha_t, sigma_t = self._sigma_to_alpha_sigma_t(sigma)
        if self.config.algorithm_type in ("dpmsolver++", "sde-dpmsolver++"):
            if pt == "epsilon":
                x0_pred = (sample - sigma_t * model_output) / alpha_t  # :40-42
            elif pt == "sample":
                x0_pred = model_output                                 # :43-44
            elif pt == "v_prediction":
                x0_pred = alpha_t * sample - sigma_t * model_output    # :45-48
            elif pt == "flow_prediction":
                x0_pred = sample - sigma * model_output                # :49-51 (sigma itself, not sigma_t)
            else:
                raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample`, "
                                 "`v_prediction`, or `flow_prediction` for the DPMSolverMultistepScheduler.")
            if self.config.thresholding:
                x0_pred = self._threshold_sample(x0_pred)              # :58-59
            return x0_pred, x0_pred                                    # intended semantics of :61 (C-1)
        if pt == "epsilon":
            epsilon = model_output                                     # :66-71
        elif pt == "sample":
            epsilon = (sample - alpha_t * model_output) / sigma_t      # :72-75
        elif pt == "v_prediction":
            epsilon = alpha_t * model_output + sigma_t * sample        # :76-79
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample`, or"
                             " `v_prediction` for the DPMSolverMultistepScheduler.")
        if self.config.thresholding:                                   # :85-90
            x0_pred = (sample - sigma_t * epsilon) / alpha_t
            x0_pred = self._threshold_sample(x0_pred)
            epsilon = (sample - alpha_t * x0_pred) / sigma_t
        x0_pred = (sample - sigma_t * epsilon) / alpha_t               # :92-94
        return epsilon, x0_pred                                        # :96

    def _lambdas(self, *sig):
        out = []
        for s in sig:
            a, st = self._sigma_to_alpha_sigma_t(s)
            out.append((a, st, torch.log(a) - torch.log(st)))
        return out

    def dpm_solver_first_order_update(self, m, sample, noise=None):
        (a_t, s_t, l_t), (a_s, s_s, l_s) = self._lambdas(self.sigmas[self.step_index + 1],
                                                          self.sigmas[self.step_index])
        h = l_t - l_s
        at = self.config.algorithm_type
        if at == "dpmsolver++":
            return (s_t / s_s) * sample - (a_t * (torch.exp(-h) - 1.0)) * m
        if at == "dpmsolver":
            return (a_t / a_s) * sample - (s_t * (torch.exp(h) - 1.0)) * m
        if at == "sde-dpmsolver++":
            return ((s_t / s_s * torch.exp(-h)) * sample + (a_t * (1 - torch.exp(-2.0 * h))) * m
                    + s_t * torch.sqrt(1.0 - torch.exp(-2.0 * h)) * noise)
        return ((a_t / a_s) * sample - 2.0 * (s_t * (torch.exp(h) - 1.0)) * m
                + s_t * torch.sqrt(torch.exp(2.0 * h) - 1.0) * noise)

    def multistep_dpm_solver_second_order_update(self, ms, sample, noise=None):
        i = self.step_index
        (a_t, s_t, l_t), (a_s0, s_s0, l_s0), (_, _, l_s1) = self._lambdas(
            self.sigmas[i + 1], self.sigmas[i], self.sigmas[i - 1])
        m0, m1 = ms[-1], ms[-2]
        h, h_0 = l_t - l_s0, l_s0 - l_s1
        r0 = h_0 / h
        D0, D1 = m0, (1.0 / r0) * (m0 - m1)
        at, st = self.config.algorithm_type, self.config.solver_type
        if at == "dpmsolver++":
            if st == "midpoint":
                return ((s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * D0
                        - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * D1)
            return ((s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * D0
                    + (a_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1)
        if at == "dpmsolver":
            if st == "midpoint":
                return ((a_t / a_s0) * sample - (s_t * (torch.exp(h) - 1.0)) * D0
                        - 0.5 * (s_t * (torch.exp(h) - 1.0)) * D1)
            return ((a_t / a_s0) * sample - (s_t * (torch.exp(h) - 1.0)) * D0
                    - (s_t * ((torch.exp(h) - 1.0) / h - 1.0)) * D1)
        if at == "sde-dpmsolver++":
            if st == "midpoint":
                return ((s_t / s_s0 * torch.exp(-h)) * sample + (a_t * (1 - torch.exp(-2.0 * h))) * D0
                        + 0.5 * (a_t * (1 - torch.exp(-2.0 * h))) * D1
                        + s_t * torch.sqrt(1.0 - torch.exp(-2.0 * h)) * noise)
            return ((s_t / s_s0 * torch.exp(-h)) * sample + (a_t * (1 - torch.exp(-2.0 * h))) * D0
                    + (a_t * ((1.0 - torch.exp(-2.0 * h)) / (-2.0 * h) + 1.0)) * D1
                    + s_t * torch.sqrt(1.0 - torch.exp(-2.0 * h)) * noise)
        if st == "midpoint":
            return ((a_t / a_s0) * sample - 2.0 * (s_t * (torch.exp(h) - 1.0)) * D0
                    - (s_t * (torch.exp(h) - 1.0)) * D1 + s_t * torch.sqrt(torch.exp(2.0 * h) - 1.0) * noise)
        return ((a_t / a_s0) * sample - 2.0 * (s_t * (torch.exp(h) - 1.0)) * D0
                - 2.0 * (s_t * ((torch.exp(h) - 1.0) / h - 1.0)) * D1
                + s_t * torch.sqrt(torch.exp(2.0 * h) - 1.0) * noise)

    def multistep_dpm_solver_third_order_update(self, ms, sample, noise=None):
        i = self.step_index
        (a_t, s_t, l_t), (a_s0, s_s0, l_s0), (_, _, l_s1), (_, _, l_s2) = self._lambdas(
            self.sigmas[i + 1], self.sigmas[i], self.sigmas[i - 1], self.sigmas[i - 2])
        m0, m1, m2 = ms[-1], ms[-2], ms[-3]
        h, h_0, h_1 = l_t - l_s0, l_s0 - l_s1, l_s1 - l_s2
        r0, r1 = h_0 / h, h_1 / h
        D0 = m0
        D1_0, D1_1 = (1.0 / r0) * (m0 - m1), (1.0 / r1) * (m1 - m2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        if self.config.algorithm_type == "dpmsolver++":
            return ((s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * D0
                    + (a_t * ((torch.exp(-h) - 1.0) / h + 1.0)) * D1
                    - (a_t * ((torch.exp(-h) - 1.0 + h) / h ** 2 - 0.5)) * D2)
        return ((a_t / a_s0) * sample - (s_t * (torch.exp(h) - 1.0)) * D0
                - (s_t * ((torch.exp(h) - 1.0) / h - 1.0)) * D1
                - (s_t * ((torch.exp(h) - 1.0 - h) / h ** 2 - 0.5)) * D2)

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict=False):
        """/root/reference/src/schedulers.py:98-187."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' "
                             "after creating the scheduler")
        if self.step_index is None:
            self._init_step_index(timestep)
        n = len(self.timesteps)
        lower_order_final = (self.step_index == n - 1) and (
            self.config.euler_at_final or (self.config.lower_order_final and n < 15)
            or self.config.final_sigmas_type == "zero")
        lower_order_second = (self.step_index == n - 2) and self.config.lower_order_final and n < 15

        model_output, x0_pred = self.convert_model_output(model_output, sample=sample)
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = model_output

        sample = sample.to(torch.float32)
        if self.config.algorithm_type in ("sde-dpmsolver", "sde-dpmsolver++") and variance_noise is None:
            noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                                 dtype=torch.float32)
        elif self.config.algorithm_type in ("sde-dpmsolver", "sde-dpmsolver++"):
            noise = variance_noise.to(device=model_output.device, dtype=torch.float32)
        else:
            noise = None

        if self.config.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = self.dpm_solver_first_order_update(model_output, sample=sample, noise=noise)
        elif self.config.solver_order == 2 or self.lower_order_nums < 2 or lower_order_second:
            prev = self.multistep_dpm_solver_second_order_update(self.model_outputs, sample=sample, noise=noise)
        else:
            prev = self.multistep_dpm_solver_third_order_update(self.model_outputs, sample=sample)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        prev = prev.to(model_output.dtype)
        self._step_index += 1
        return (prev, x0_pred)


# ----------------------------------------------------------------------------- LCM


class LCMScheduler(_SchedulerBase):
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
        trained_betas=None, original_inference_steps=50, clip_sample=False, set_alpha_to_one=True,
        steps_offset=0, prediction_type="epsilon", thresholding=False, timestep_spacing="leading",
        timestep_scaling=10.0,
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.sigma_data = 0.5

    def set_timesteps(self, num_inference_steps, device=None, **_):
        T = self.config.num_train_timesteps
        orig = self.config.original_inference_steps
        if num_inference_steps > T or num_inference_steps > orig:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`original_inference_steps`: {orig}")
        k = T // orig
        origin = np.asarray(list(range(1, int(orig * 1.0) + 1))) * k - 1
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(origin[idx]).to(device=device, dtype=torch.long)
        self._step_index = None

    def get_scalings_for_boundary_condition_discrete(self, timestep):
        scaled = timestep * self.config.timestep_scaling
        c_skip = self.sigma_data ** 2 / (scaled ** 2 + self.sigma_data ** 2)
        c_out = scaled / (scaled ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out

    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        if self.step_index is None:
            self._init_step_index(timestep)
        nxt = self.step_index + 1
        prev_t = self.timesteps[nxt] if nxt < len(self.timesteps) else timestep
        a = self.alphas_cumprod[int(timestep)]
        ap = self.alphas_cumprod[int(prev_t)] if int(prev_t) >= 0 else self.final_alpha_cumprod
        beta, beta_prev = 1 - a, 1 - ap
        c_skip, c_out = self.get_scalings_for_boundary_condition_discrete(
            torch.as_tensor(timestep).to("cpu"))
        pt = self.config.prediction_type                   # diffusers 0.32.1 scheduling_lcm.py ``step`` item 3
        if pt == "epsilon":
            x0 = (sample - beta.sqrt() * model_output) / a.sqrt()
        elif pt == "sample":
            x0 = model_output
        elif pt == "v_prediction":
            x0 = a.sqrt() * sample - beta.sqrt() * model_output
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, `sample` or `v_prediction`")
        if self.config.clip_sample:
            x0 = x0.clamp(-1.0, 1.0)
        denoised = c_out * x0 + c_skip * sample
        if self.step_index != self.num_inference_steps - 1:
            z = randn_tensor(model_output.shape, generator=generator, device=model_output.device,
                             dtype=denoised.dtype)
            prev = ap.sqrt() * denoised + beta_prev.sqrt() * z
        else:
            prev = denoised
        self._step_index += 1
        return (prev, denoised)


# ----------------------------------------------------------------------------- PNDM (PLMS)


class PNDMScheduler(_SchedulerBase):
    _defaults = dict(
        num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
        trained_betas=None, skip_prk_steps=False, set_alpha_to_one=False, prediction_type="epsilon",
        timestep_spacing="leading", steps_offset=0,
    )

    def __init__(self, **kw):
        super().__init__(**kw)
        assert self.config.skip_prk_steps, "only the PLMS path (skip_prk_steps=True) is on the reference path"
        self.final_alpha_cumprod = torch.tensor(1.0) if self.config.set_alpha_to_one else self.alphas_cumprod[0]
        self.ets, self.counter, self.cur_sample = [], 0, None

    def set_timesteps(self, num_inference_steps, device=None, **_):
        T = self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        ratio = T // num_inference_steps
        _t = (np.arange(0, num_inference_steps) * ratio).round()
        _t += self.config.steps_offset
        plms = np.concatenate([_t[:-1], _t[-2:-1], _t[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64)).to(device)
        self.ets, self.counter, self.cur_sample = [], 0, None

    def step(self, model_output, timestep, sample, return_dict=False):
        t = int(timestep)
        stride = self.config.num_train_timesteps // self.num_inference_steps
        prev_t = t - stride
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = t
            t = t + stride
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3]
                                       - 9 * self.ets[-4])
        prev = self._get_prev_sample(sample, t, prev_t, model_output)
        self.counter += 1
        return (prev,)

    def _get_prev_sample(self, sample, t, prev_t, model_output):
        a = self.alphas_cumprod[t]
        ap = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta, beta_prev = 1 - a, 1 - ap
        if self.config.prediction_type == "v_prediction":  # the (combined) v outputs become a noise prediction here
            model_output = (a ** 0.5) * model_output + (beta ** 0.5) * sample
        elif self.config.prediction_type != "epsilon":
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon` or "
                             "`v_prediction`")
        sample_coeff = (ap / a) ** 0.5
        denom = a * beta_prev ** 0.5 + (a * beta * ap) ** 0.5
        return sample_coeff * sample - (ap - a) * model_output / denom


def sd15(cls, **overrides):
    """Build a scheduler the way base_experiment.py:66-72 does: from the SD-v1.5 PNDM config."""
    return cls.from_config(SD15_SCHEDULER_CONFIG, **overrides)
