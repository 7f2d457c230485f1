"""The oracle (and the product's host logic) against fixtures produced by EXECUTING THE REFERENCE'S OWN SOURCE.

``tests/golden/reference_pins.{npz,json}`` were written by ``tests/golden/make_reference_pins.py``, which
compiles /root/reference/src/{schedulers,models,registry}.py + utils/class_registry.py where they lie
(oracle/refexec.py) and runs the seeded cases of ``tests/refpin_cases.py`` through them.  Here:

  * wherever ``/root/reference`` exists (the build container) the fixtures are re-derived live and must be
    BIT-IDENTICAL to the committed files, and the oracle must be bit-identical to the reference source;
  * everywhere (the GPU box has no reference tree) the oracle is compared with the committed fixtures:
    integer schedules exactly, float tensors to 1e-5 of their range (another host's libm / conv summation
    order may differ in the last bits).
"""
import dataclasses
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import refpin_cases as RC  # noqa: E402
from oracle import refexec  # noqa: E402
from oracle import schedulers as O  # noqa: E402

PINS = np.load(os.path.join(HERE, "golden", "reference_pins.npz"))
META = json.load(open(os.path.join(HERE, "golden", "reference_pins.json")))
LIVE = refexec.available()
needs_reference = pytest.mark.skipif(not LIVE, reason="/root/reference is absent (GPU box): fixtures only")


@pytest.fixture(scope="module", autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)              # the fixtures were generated single-threaded
    yield
    torch.set_num_threads(n)


@pytest.fixture(scope="module")
def net():
    return RC.tiny_unet()


def _close(got, want, what):
    want = torch.from_numpy(np.asarray(want))
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if LIVE and torch.equal(got, want):
        return
    scale = max(1.0, want.abs().max().item())
    err = (got - want).abs().max().item() / scale
    assert err <= 1e-5, f"{what}: {err:.3e} of range"


# --------------------------------------------------------------------------- scheduler step (schedulers.py:14-187)
ALL_SCHEDULER_CASES = RC.ALL_SCHEDULER_CASES


@pytest.mark.parametrize("name", list(ALL_SCHEDULER_CASES))
def test_oracle_scheduler_step_equals_reference_source(name):
    kind, over, n, patch, seed = ALL_SCHEDULER_CASES[name]
    prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=O), n, seed)
    assert ts == META["scheduler_timesteps"][name]
    _close(torch.stack(prevs), PINS[f"sched/{name}/prev"], f"{name} prev_sample")
    _close(torch.stack(x0s), PINS[f"sched/{name}/x0"], f"{name} x0_pred")


@needs_reference
@pytest.mark.parametrize("name", list(ALL_SCHEDULER_CASES))
def test_scheduler_fixtures_are_what_the_reference_source_computes(name):
    kind, over, n, patch, seed = ALL_SCHEDULER_CASES[name]
    ns = refexec.load(patch_c1=patch)
    prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, ref=ns), n, seed)
    assert ts == META["scheduler_timesteps"][name]
    assert np.array_equal(torch.stack(prevs).numpy(), PINS[f"sched/{name}/prev"])
    assert np.array_equal(torch.stack(x0s).numpy(), PINS[f"sched/{name}/x0"])
    # ... and the oracle is bit-identical to it on this host
    po, xo, _ = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=O), n, seed)
    assert all(torch.equal(a, b) for a, b in zip(po, prevs)) and all(torch.equal(a, b) for a, b in zip(xo, x0s))


@pytest.mark.parametrize("name", list(ALL_SCHEDULER_CASES))
def test_product_coefficients_reproduce_reference_source(name, monkeypatch):
    """The product's reduction of ``convert_model_output`` + ``step`` to the fused kernel's linear-combination
    coefficients (every ``prediction_type`` and solver order), with the kernel emulated in float64 on the CPU,
    against what the reference's own source computed.  The same cases run through the real kernel in
    tests/test_reference_pins_gpu.py."""
    from test_host_cpu import _emulated_launch

    from sonicdiffusionbayeslab_b200 import schedulers as S

    monkeypatch.setattr(S.FusedScheduler, "_launch", _emulated_launch)
    kind, over, n, patch, seed = ALL_SCHEDULER_CASES[name]
    want_prev = torch.from_numpy(PINS[f"sched/{name}/prev"])
    want_x0 = torch.from_numpy(PINS[f"sched/{name}/x0"])
    prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=S), n, seed)   # free-running
    assert ts == META["scheduler_timesteps"][name]
    for got, want in list(zip(prevs, want_prev)) + list(zip(x0s, want_x0)):
        assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item()), name


def _emulated_launch_model_dtype(self, coeffs, eps, eps_text, sample, hist=(), noise=None, want_m0=False, want_x0=True,
                                 out=None, ring=True, post=None):
    """Model of ``latent_update_kernel`` (csrc/elementwise.cu ``update_math``) with its real arithmetic: fp32 math,
    the guided model output and the converted output rounded to the I/O dtype, one rounding per stored tensor."""
    f, dt = torch.float32, sample.dtype
    c = {k: float(np.float32(coeffs.get(k, 0.0))) for k in ("guidance", "m_x", "m_e", "x0_x", "x0_e", "c_x", "c_e",
                                                            "c_m0", "c_h1", "c_h2", "c_h3", "c_z")}
    e = eps.to(f)
    if eps_text is not None:
        e = e + c["guidance"] * (eps_text.to(f) - e)
    e, x = e.to(dt).to(f), sample.to(f)
    m0 = (c["m_x"] * x + c["m_e"] * e).to(dt).to(f)
    x0 = c["x0_x"] * x + c["x0_e"] * e
    if post is not None:                             # latent_update_post_kernel: x0 rounded, processed, m0 re-derived
        from test_host_cpu import _emulated_post

        x0, m0 = _emulated_post(x0.to(dt).to(f), x, post)
        m0 = m0.to(dt).to(f)
    h = [t.to(f) for t in hist] + [torch.zeros_like(x)] * (3 - len(hist))
    z = torch.zeros_like(x) if noise is None else noise.to(f)
    xn = (c["c_x"] * x + c["c_e"] * e + c["c_m0"] * m0 + c["c_h1"] * h[0] + c["c_h2"] * h[1] + c["c_h3"] * h[2]
          + c["c_z"] * z)
    return xn.to(dt), (m0.to(dt) if want_m0 else None), (x0.to(dt) if want_x0 else None)


@pytest.mark.parametrize("name", list(ALL_SCHEDULER_CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kernel_arithmetic_model_meets_the_gpu_tolerances(name, dtype, monkeypatch):
    """tests/test_reference_pins_gpu.py on the CPU with the kernel replaced by its arithmetic model: the stated
    tolerances (refpin_cases.step_tolerance: fp32 <= 1e-4, bf16 <= 2e-2 of the range, teacher-forced) hold for the
    kernel's rounding points with a 2.4x margin (1.6x for the thresholded bf16 cases), so a GPU failure there is a
    kernel bug, not a tolerance that was set by looking at the GPU."""
    from sonicdiffusionbayeslab_b200 import schedulers as S

    monkeypatch.setattr(S.FusedScheduler, "_launch", _emulated_launch_model_dtype)
    kind, over, n, patch, seed = ALL_SCHEDULER_CASES[name]
    want_prev = torch.from_numpy(PINS[f"sched/{name}/prev"])
    want_x0 = torch.from_numpy(PINS[f"sched/{name}/x0"])
    prevs, x0s, _ = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=S), n, seed, dtype=dtype,
                                          teacher=list(want_prev))
    worst = 0.0
    for got, want in list(zip(prevs, want_prev)) + list(zip(x0s, want_x0)):
        scale = max(1.0, want.abs().max().item()) if dtype == torch.bfloat16 else 1.0
        worst = max(worst, (got.float() - want).abs().max().item() / scale)
    margin = 1.6 if (name in RC.THRESHOLD_CASES and dtype == torch.bfloat16) else 2.4
    assert worst <= RC.step_tolerance(name, dtype) / margin, (name, dtype, worst)


def test_product_scheduler_prediction_type_errors():
    """Error behaviour of src/schedulers.py:52-56 / :80-83 (ValueError for an unknown ``prediction_type``;
    ``flow_prediction`` exists only for the ``++`` algorithms)."""
    from sonicdiffusionbayeslab_b200 import schedulers as S

    with pytest.raises(ValueError):
        S.DPMSolverScheduler.from_config(RC.SD15, prediction_type="nonsense")
    with pytest.raises(ValueError):
        S.DPMSolverScheduler.from_config(RC.SD15, algorithm_type="dpmsolver", final_sigmas_type="sigma_min",
                                         prediction_type="flow_prediction")
    S.DPMSolverScheduler.from_config(RC.SD15, prediction_type="flow_prediction")
    with pytest.raises(NotImplementedError):
        S.DPMSolverScheduler.from_config(RC.SD15, use_karras_sigmas=True)


# --------------------------------------------------------------------------- pipelines (models.py call bodies)
@pytest.mark.parametrize("name", list(RC.PIPELINE_CASES))
def test_oracle_pipeline_equals_reference_source(name, net):
    r = RC.run_pipeline_oracle(RC.PIPELINE_CASES[name], net)
    assert r["timesteps"] == META["pipeline_timesteps"][name]
    _close(torch.stack(r["per_step"]), PINS[f"pipe/{name}/per_step"], f"{name} per-step latents")


@needs_reference
@pytest.mark.parametrize("name", list(RC.PIPELINE_CASES))
def test_pipeline_fixtures_are_what_the_reference_source_computes(name, net):
    case = RC.PIPELINE_CASES[name]
    r = RC.run_pipeline_reference(case, refexec.load(patch_c1=case["patch"]), net)
    assert r["timesteps"] == META["pipeline_timesteps"][name]
    assert np.array_equal(torch.stack(r["per_step"]).numpy(), PINS[f"pipe/{name}/per_step"])
    assert {"n_x0": r["n_x0"], "num_timesteps": r["num_timesteps"]} == META["pipeline_info"][name]
    o = RC.run_pipeline_oracle(case, net)
    assert all(torch.equal(a, b) for a, b in zip(o["per_step"], r["per_step"]))      # oracle == reference source


def test_product_host_schedules_equal_reference_source():
    """The product's timestep / index logic (host side, no GPU needed) on the pinned pipeline cases:
    bit-exact integer lists, as BASELINE.json's north_star requires."""
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    for name, case in RC.PIPELINE_CASES.items():
        want = META["pipeline_timesteps"][name]
        kind = case["pipe"]
        if kind in ("single", "skip"):
            s = RC.make_scheduler(*case["sched"], module=S)
            s.set_timesteps(case["steps"])
            ts = [int(t) for t in s.timesteps.tolist()]
            if kind == "skip":
                ts = [t for i, t in enumerate(ts) if i not in set(case["skip"])]
        elif kind == "two":
            s1, s2 = RC.make_scheduler(*case["first"], module=S), RC.make_scheduler(*case["second"], module=S)
            ts1, _ = M.retrieve_timesteps(s1, case["n1"], None, None)
            ts2, _ = M.retrieve_timesteps(s2, device=None, timesteps=ts1.cpu().numpy())
            first, second = M.StableDiffusionModelTwoSchedulers.switch_timestamp(None, ts1, ts2, case["k"],
                                                                                 case["type_switch"])
            ts = [int(t) for t in first + second]
        else:
            main = RC.make_scheduler(*case["main"], module=S)
            main.set_timesteps(case["steps"])
            ts, _ = M.StableDiffusionModelInterlivingSchedulers.partition(main.timesteps.tolist(),
                                                                          main.config.solver_order, case["groups"])
        assert ts == want, name


# --------------------------------------------------------------------------- switch_timestamp (models.py:704-730)
def test_switch_timestamp_equals_reference_source():
    from oracle.pipeline import switch_timestamp
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    live = refexec.load().StableDiffusionModelTwoSchedulers.switch_timestamp if LIVE else None
    assert len(META["switch"]) == len(RC.SWITCH_CASES) * len(RC.SWITCH_TYPES)
    for e in META["switch"]:
        def grids(mod):
            s1 = (getattr(mod, "DDIMSchedulerMy", None) or mod.DDIMScheduler).from_config(RC.SD15)
            s1.set_timesteps(e["n1"])
            s2 = mod.DPMSolverScheduler.from_config(RC.SD15)
            if e["n2"] is None:
                s2.set_timesteps(timesteps=s1.timesteps.cpu().numpy())
            else:
                s2.set_timesteps(e["n2"])
            return s1.timesteps, s2.timesteps

        impls = [("oracle", lambda a, b: switch_timestamp(a, b, e["k"], e["type_switch"]), O),
                 ("product", lambda a, b: M.StableDiffusionModelTwoSchedulers.switch_timestamp(
                     None, a, b, e["k"], e["type_switch"]), S)]
        if live is not None:
            impls.append(("reference", lambda a, b: live(None, a, b, e["k"], e["type_switch"]), O))
        for label, fn, mod in impls:
            a, b = grids(mod)
            if "raises" in e:
                with pytest.raises(IndexError):
                    fn(a, b)
                continue
            first, second = fn(a, b)
            assert [int(t) for t in first] == e["first"] and [int(t) for t in second] == e["second"], (label, e)
            assert all(isinstance(t, np.integer) for t in list(first) + list(second)), label   # lists of np.int64


# --------------------------------------------------------------------------- the defects the product does not reproduce
def test_reference_defects_are_pinned():
    want = {"c1_dpmpp_step_batch3": "ValueError", "c4_two_ddim_dpm_unpatched": "RuntimeError",
            "c4_two_ddim_dpm_c1_patched": "RuntimeError", "c15_two_pndm_default_solver_order": "KeyError"}
    assert META["raises"] == want
    if not LIVE:
        return
    ref = refexec.load()
    with pytest.raises(ValueError):            # SURVEY C-1: ``a, b = tensor`` with a batch of 3
        RC.run_scheduler_case(RC.make_scheduler("dpm", dict(algorithm_type="dpmsolver++"), ref=ref), 5, None)


# --------------------------------------------------------------------------- registry (class_registry.py:17-68)
def test_registry_matches_reference_source():
    from sonicdiffusionbayeslab_b200 import registry as R
    from sonicdiffusionbayeslab_b200.utils.class_registry import ClassRegistry

    import sonicdiffusionbayeslab_b200.models  # noqa: F401  (registration is an import side effect)
    import sonicdiffusionbayeslab_b200.schedulers  # noqa: F401

    class Probe:
        def __init__(self, a, b=None, c=3, d="x", e=1.5, *args, **kwargs):
            pass

    reg = ClassRegistry()
    reg.add_to_registry("probe")(Probe)
    got = [[f.name, str(f.type), None if f.default is dataclasses.MISSING else repr(f.default)]
           for f in dataclasses.fields(reg.args["probe"])]
    assert got == META["registry"]["probe_fields"]
    assert reg["probe"] is Probe
    names = META["registry"]["reference_names"]
    assert set(names["models"]) <= set(R.models_registry.classes)
    assert set(names["schedulers"]) <= set(R.schedulers_registry.classes)
