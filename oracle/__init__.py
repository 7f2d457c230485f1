"""CPU restatement oracle of the SonicDiffusionBayesLab hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sonicdiffusionbayeslab_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only
as the checker / the timed CPU baseline, never as the product.

PARITY PIN (round 2): the reference ships no tests, golden vectors or known-answer
fixtures, and the arithmetic UNDER its code lives in third-party
packages (diffusers 0.32.1, DeepCache 0.1.1, torchmetrics 1.6.1; pins at
/root/reference/poetry.lock:436-455,2652-2653) that are neither vendored under
/root/reference nor installable in this image -- so that layer is restated from the
published algorithms and stays UNPINNED (UNet, the DDIM / LCM / PNDM formulas, the
DPM-Solver update formulas, DeepCache, CLIP score).  What the reference itself OWNS on
the hot path IS pinned, by executing its source files where they lie
(``oracle/refexec.py``; fixtures in tests/golden/reference_pins.{npz,json} written by
tests/golden/make_reference_pins.py, checked by tests/test_reference_pins_{cpu,gpu}.py):

  * denoising loop            /root/reference/src/models.py:21-335 (bit-identical to ``pipeline.denoise``)
  * two-scheduler call/switch /root/reference/src/models.py:338-730
  * interleaved schedulers    /root/reference/src/models.py:733-1135
  * skip-timesteps loop       /root/reference/src/models.py:1138-1467
  * DPM-Solver override       /root/reference/src/schedulers.py:14-187 (``convert_model_output`` + ``step``)
  * plugin registry           /root/reference/src/utils/class_registry.py:8-68, src/registry.py:3-6
  * experiment drivers        /root/reference/src/experiments/*.py (all eight methods, over the recording fake
                              backend of tests/driver_cases.py; tests/test_reference_drivers_cpu.py)
  * metric plugins / dataset  /root/reference/src/metrics/metrics.py:25-41,115-131 (``calc_metric`` batching,
                              ``TimeMetric``), src/dataset/dataset.py, src/utils/model_utils.py
                              (``refexec.load_plugins``; tests/test_reference_plugins_cpu.py)

Restated from call sites only (nothing executable without the absent packages):

  * DDIM / LCM subclasses     /root/reference/src/schedulers.py:190-197 (empty subclasses of diffusers)
  * DeepCache call sites      /root/reference/src/experiments/deep_cache.py:24-29,58
  * CLIP score                /root/reference/src/metrics/metrics.py:25-41 (towers pinned vs ``transformers``)
  * generator / dtype / args  /root/reference/src/experiments/base_experiment.py:51-72,122-163

plus the closed-form known answers of SURVEY.md appendix A.6
(``tests/golden/schedule_kat.json``), re-derived independently in
``tests/golden/make_golden.py``.
"""
