"""The product's experiment drivers (sonicdiffusionbayeslab_b200/experiments) against the reference's OWN drivers
(/root/reference/src/experiments/*.py, executed from source over stub modules): both are run over the same recording
fake backend (tests/driver_cases.py) and must leave the same event log -- scheduler construction (class + solver keys),
LoRA and DeepCache helper calls, every call into the model plugin with its keyword arguments, image-log names (incl.
the "X0 preds" batches of default_sd.py / skip_steps_exp.py), and every metric table with its extra columns and the
point at which the table is started afresh.

Wherever ``/root/reference`` exists the reference logs are re-derived live and must equal the committed
``tests/golden/reference_driver_events.json``; everywhere the product is compared with the committed logs.
"""
import contextlib
import io
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import driver_cases as DC  # noqa: E402
from oracle import refexec  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "reference_driver_events.json")))
LIVE = refexec.available()


def _quiet(fn, *a):
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


@pytest.mark.parametrize("name", list(DC.DRIVER_CASES))
def test_product_driver_leaves_the_reference_drivers_event_log(name, monkeypatch):
    want = GOLD["events"][name]
    if LIVE:                                               # the fixture is what the reference source does today
        assert json.loads(json.dumps(_quiet(DC.run_reference_driver, DC.DRIVER_CASES[name]))) == want
    got = json.loads(json.dumps(_quiet(DC.run_product_driver, DC.DRIVER_CASES[name], monkeypatch)))
    assert DC.normalise(got) == DC.normalise(want)
    # C-9 (the one deliberate behavioural difference): the reference counts a ragged last batch as a full one
    n_calls = sum(e[0] == "call" for e in want)
    assert DC.time_updates(want) == [DC.BATCH] * n_calls
    assert DC.time_updates(got) == [e[1] for e in got if e[0] == "call"]


def test_reference_driver_defects_are_pinned(monkeypatch):
    """SURVEY C-3: with the solver keys of the shipped two_schedulers_config.yaml left unset the reference forwards
    ``algorithm_type=""`` and the DPM-Solver constructor refuses it; the product reads "" as "class default"."""
    assert GOLD["raises"] == {"two_schedulers_shipped_keys": "NotImplementedError"}
    case = DC.RAISING_CASES["two_schedulers_shipped_keys"]
    if LIVE:
        with pytest.raises(NotImplementedError):
            _quiet(DC.run_reference_driver, case)
    events = _quiet(DC.run_product_driver, case, monkeypatch)
    second = [e for e in events if e[0] == "scheduler" and e[1] == "scheduler_second"][0]
    assert second[2] == "DPMSolverScheduler" and second[3]["algorithm_type"] == "dpmsolver++"
