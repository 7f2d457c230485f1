"""CPU restatement oracle of the SonicDiffusionBayesLab hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sonicdiffusionbayeslab_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only
as the checker / the timed CPU baseline, never as the product.

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer
fixtures, and its arithmetic lives in third-party
packages (diffusers 0.32.1, DeepCache 0.1.1, torchmetrics 1.6.1; pins at
/root/reference/poetry.lock:436-455,2652-2653) that are neither vendored under
/root/reference nor installable in this image.  This package therefore
restates the *published algorithms* of those packages in plain PyTorch
(fp32, CPU) and anchors on the reference's own call sites:

  * denoising loop            /root/reference/src/models.py:210-282
  * two-scheduler switch      /root/reference/src/models.py:487-502,545-621,704-730
  * DPM-Solver override       /root/reference/src/schedulers.py:14-187
  * DDIM / LCM subclasses     /root/reference/src/schedulers.py:190-197
  * DeepCache call sites      /root/reference/src/experiments/deep_cache.py:24-29,58
  * CLIP score                /root/reference/src/metrics/metrics.py:25-41
  * generator / dtype / args  /root/reference/src/experiments/base_experiment.py:51-72,122-163

The only pins are the closed-form known answers of SURVEY.md appendix A.6
(``tests/golden/schedule_kat.json``), re-derived independently in
``tests/golden/make_golden.py``.
"""
