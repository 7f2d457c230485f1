#!/usr/bin/env python
"""Generates ``tests/golden/reference_pins.npz`` + ``reference_pins.json`` by EXECUTING THE REFERENCE'S OWN
SOURCE (oracle/refexec.py compiles /root/reference/src/{schedulers,models,registry}.py where they lie, over
stub third-party modules) on the seeded cases of ``tests/refpin_cases.py``.

Run in the build container (``/root/reference`` is absent on the GPU box):

    python tests/golden/make_reference_pins.py

What each fixture pins
  sched/<case>/{prev,x0}   /root/reference/src/schedulers.py:14-187 ``convert_model_output`` + ``step`` (order
                           selection, lower-order-final, history shift, fp32 upcast, return tuple) over the
                           synthetic epsilon sequence; the ``++`` cases run with the one-line C-1 source patch
                           (oracle/refexec.py ``C1_PATCH``), the others unmodified; every ``prediction_type``
                           branch (:36-56, :65-83) and -- for the oracle only -- the thresholding branches
                           (:58-59, :85-90) over the oracle's restatement of diffusers' ``_threshold_sample``
  pipe/<case>/per_step     /root/reference/src/models.py ``call`` of the four pipeline classes (loop body,
                           CFG combine, RNG order, two-scheduler switch, interleave partition, skip mask)
                           over a tiny oracle UNet
  json: switch             ``switch_timestamp`` (models.py:704-730) for all three ``type_switch`` modes
  json: raises             the defects the product deliberately does not reproduce: C-1 (``++`` step raises
                           for B != 2), C-4 (DDIM -> DPM two-scheduler call raises in the history seeding) and
                           the ``self.scheduler.config.solver_order`` read of models.py:638 with the stock PNDM
                           default scheduler
  json: registry           ``ClassRegistry.add_to_registry`` argument dataclasses (class_registry.py:17-68)
  reference_driver_events.json   the event logs of the reference's OWN experiment drivers (src/experiments/*.py, all
                           eight methods) over the recording fake backend of tests/driver_cases.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import driver_cases as DC  # noqa: E402
import refpin_cases as RC  # noqa: E402
from oracle import refexec  # noqa: E402
from oracle import schedulers as O  # noqa: E402


def main():
    torch.set_num_threads(1)                       # one summation order, whatever the host
    ref = refexec.load(patch_c1=False)
    ref_c1 = refexec.load(patch_c1=True)
    arrays, meta = {}, {"scheduler_timesteps": {}, "pipeline_timesteps": {}, "pipeline_info": {}, "switch": [],
                        "raises": {}, "registry": {}, "callback_shapes": {}, "legacy_callback": {}}

    for name, (kind, over, n, patch, seed) in {**RC.SCHEDULER_CASES, **RC.THRESHOLD_CASES}.items():
        ns = ref_c1 if patch else ref
        prevs, x0s, ts = RC.run_scheduler_case(RC.make_scheduler(kind, over, ref=ns), n, seed)
        arrays[f"sched/{name}/prev"] = torch.stack(prevs).numpy()
        arrays[f"sched/{name}/x0"] = torch.stack(x0s).numpy()
        meta["scheduler_timesteps"][name] = ts

    net = RC.tiny_unet()
    for name, case in RC.PIPELINE_CASES.items():
        ns = ref_c1 if case["patch"] else ref
        r = RC.run_pipeline_reference(case, ns, net)
        arrays[f"pipe/{name}/per_step"] = torch.stack(r["per_step"]).numpy()
        meta["pipeline_timesteps"][name] = r["timesteps"]
        meta["pipeline_info"][name] = {"n_x0": r["n_x0"], "num_timesteps": r["num_timesteps"]}
        assert torch.equal(r["final"], r["per_step"][-1])
        if r["cb_shapes"] is not None:                 # what ``callback_on_step_end`` is handed (models.py:263-267)
            meta["callback_shapes"][name] = [{k: list(v) for k, v in d.items()} for d in r["cb_shapes"]]
        if r["legacy_calls"] is not None:              # the deprecated ``callback`` (models.py:275-282)
            meta["legacy_callback"][name] = r["legacy_calls"]

    # switch_timestamp, called unbound from the compiled reference class (it does not touch ``self``)
    switch = ref.StableDiffusionModelTwoSchedulers.switch_timestamp
    for n1, n2, k in RC.SWITCH_CASES:
        s1 = O.DDIMScheduler.from_config(RC.SD15)
        s1.set_timesteps(n1)
        s2 = O.DPMSolverScheduler.from_config(RC.SD15)
        if n2 is None:
            s2.set_timesteps(timesteps=s1.timesteps.cpu().numpy())
        else:
            s2.set_timesteps(n2)
        for mode in RC.SWITCH_TYPES:
            entry = {"n1": n1, "n2": n2, "k": k, "type_switch": mode}
            try:
                first, second = switch(None, s1.timesteps, s2.timesteps, k, mode)
                entry.update(first=[int(t) for t in first], second=[int(t) for t in second])
            except IndexError:                     # no candidate on that side of the pivot (models.py:719,728)
                entry["raises"] = "IndexError"
            meta["switch"].append(entry)

    # defects: what the UNPATCHED reference does
    def raised(fn):
        try:
            fn()
        except Exception as e:                     # noqa: BLE001
            return type(e).__name__
        return None

    meta["raises"]["c1_dpmpp_step_batch3"] = raised(lambda: RC.run_scheduler_case(
        RC.make_scheduler("dpm", dict(algorithm_type="dpmsolver++"), ref=ref), 5, None))
    c4 = dict(pipe="two", first=("ddim", {}), second=("dpm", dict(algorithm_type="dpmsolver++")), n1=10, k=3,
              type_switch="closest", guidance=7.5)
    c15 = dict(RC.PIPELINE_CASES["two_ddim_dpmstock"], pndm_default=True)
    meta["raises"]["c15_two_pndm_default_solver_order"] = raised(lambda: RC.run_pipeline_reference(c15, ref, net))
    meta["raises"]["c4_two_ddim_dpm_unpatched"] = raised(lambda: RC.run_pipeline_reference(c4, ref, net))
    meta["raises"]["c4_two_ddim_dpm_c1_patched"] = raised(lambda: RC.run_pipeline_reference(c4, ref_c1, net))

    # registry: argument dataclasses the reference's ClassRegistry derives from an __init__ signature
    reg = ref.ClassRegistry()

    class Probe:
        def __init__(self, a, b=None, c=3, d="x", e=1.5, *args, **kwargs):
            pass

    reg.add_to_registry("probe")(Probe)
    import dataclasses

    meta["registry"]["probe_fields"] = [[f.name, str(f.type), None if f.default is dataclasses.MISSING else
                                         repr(f.default)] for f in dataclasses.fields(reg.args["probe"])]
    meta["registry"]["reference_names"] = {
        "models": sorted(ref.registry.models_registry.classes), "schedulers": sorted(ref.registry.schedulers_registry.classes)}

    # the reference's own experiment drivers (src/experiments/*.py) over the recording fake backend
    drivers = {"events": {}, "raises": {}}
    import contextlib
    import io

    for name, case in DC.DRIVER_CASES.items():
        with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
            drivers["events"][name] = DC.run_reference_driver(case)
    for name, case in DC.RAISING_CASES.items():
        try:
            with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
                DC.run_reference_driver(case)
            drivers["raises"][name] = None
        except Exception as e:                            # noqa: BLE001
            drivers["raises"][name] = type(e).__name__
    with open(os.path.join(HERE, "reference_driver_events.json"), "w") as f:
        json.dump(drivers, f, indent=1, sort_keys=True)
    print(f"wrote {len(drivers['events'])} driver event logs, driver raises = {drivers['raises']}")

    np.savez_compressed(os.path.join(HERE, "reference_pins.npz"), **arrays)
    with open(os.path.join(HERE, "reference_pins.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    size = os.path.getsize(os.path.join(HERE, "reference_pins.npz"))
    print(f"wrote {len(arrays)} arrays ({size / 1e6:.2f} MB), raises = {meta['raises']}")


if __name__ == "__main__":
    main()
