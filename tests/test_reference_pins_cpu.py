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
@pytest.mark.parametrize("name", list(RC.SCHEDULER_CASES))
def test_oracle_scheduler_step_equals_reference_source(name):
    kind, over, n, patch, seed = RC.SCHEDULER_CASES[name]
    prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=O), n, seed)
    assert ts == META["scheduler_timesteps"][name]
    _close(torch.stack(prevs), PINS[f"sched/{name}/prev"], f"{name} prev_sample")
    _close(torch.stack(x0s), PINS[f"sched/{name}/x0"], f"{name} x0_pred")


@needs_reference
@pytest.mark.parametrize("name", list(RC.SCHEDULER_CASES))
def test_scheduler_fixtures_are_what_the_reference_source_computes(name):
    kind, over, n, patch, seed = RC.SCHEDULER_CASES[name]
    ns = refexec.load(patch_c1=patch)
    prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, ref=ns), n, seed)
    assert ts == META["scheduler_timesteps"][name]
    assert np.array_equal(torch.stack(prevs).numpy(), PINS[f"sched/{name}/prev"])
    assert np.array_equal(torch.stack(x0s).numpy(), PINS[f"sched/{name}/x0"])
    # ... and the oracle is bit-identical to it on this host
    po, xo, _ = RC.run_scheduler_case(RC.make_scheduler(kind, over, module=O), n, seed)
    assert all(torch.equal(a, b) for a, b in zip(po, prevs)) and all(torch.equal(a, b) for a, b in zip(xo, x0s))


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
