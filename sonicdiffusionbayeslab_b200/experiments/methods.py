"""The registered methods of the reference, one thin ``BaseMethod`` subclass each:

``ddim``              /root/reference/src/experiments/ddim.py:11
``dpm_solver``        /root/reference/src/experiments/dpm_solver.py:9
``deep_cache``        /root/reference/src/experiments/deep_cache.py:10
``consistency_model`` /root/reference/src/experiments/consistency_model.py:9
``two_schedulers``    /root/reference/src/experiments/two_schedulers.py:10
``default``           /root/reference/src/experiments/default_sd.py:10
``skip_steps``        /root/reference/src/experiments/skip_steps_exp.py:10
``interliving_schedulers`` /root/reference/src/experiments/interliving_exp.py:10

They read the same ``experiment_params`` keys and build schedulers with the same
``schedulers_registry[name].from_config(model.scheduler.config, **kw)`` call.
"""
from __future__ import annotations

from ..deepcache import DeepCacheSDHelper
from ..registry import methods_registry, schedulers_registry
from .base_experiment import BaseMethod


@methods_registry.add_to_registry("ddim")
class DDIMMethod(BaseMethod):
    def setup_exp_params(self):
        self.num_inference_steps = self.config.experiment_params.num_inference_steps

    def run_experiment(self):
        bs = self.config.inference.get("batch_size", 1)
        self._new_table()
        for steps in self.num_inference_steps:
            self._sweep_point(bs, steps, f"{self.config.experiment_name}, Inference steps: {steps}", x0_grids=True)


@methods_registry.add_to_registry("dpm_solver")
class DPMSolverMethod(BaseMethod):
    def setup_exp_params(self):
        p = self.config.experiment_params
        self.num_inference_steps = p.num_inference_steps
        self.solver_order = p.solver_order
        self.algorithm_type = p.algorithm_type          # defaults supplied by config.DEFAULTS (C-2)
        self.final_sigmas_type = p.final_sigmas_type
        self.batch_size = self.config.inference.get("batch_size", 1)

    def setup_scheduler(self, **kwargs):
        return super().setup_scheduler(solver_order=self.solver_order, algorithm_type=self.algorithm_type,
                                       final_sigmas_type=self.final_sigmas_type)

    def run_experiment(self):
        self._new_table()
        for steps in self.num_inference_steps:
            self._sweep_point(self.batch_size, steps, f"{self.config.experiment_name}, Solver order: "
                                                      f"{self.solver_order}, Inference steps: {steps}", x0_grids=True)


@methods_registry.add_to_registry("consistency_model")
class ConsistencyModelMethod(BaseMethod):
    def setup_exp_params(self):
        p = self.config.experiment_params
        self.num_inference_steps = p.num_inference_steps
        self.guidance_scale = p.guidance_scale
        self.batch_size = self.config.inference.get("batch_size", 1)

    def run_experiment(self):
        self.model.load_lora_weights(self.config.experiment_params.adapter_id)
        self.model.fuse_lora()
        self._new_table()
        for steps in self.num_inference_steps:
            self._sweep_point(self.batch_size, steps, f"{self.config.experiment_name}, Inference steps: {steps}",
                              guidance_scale=self.guidance_scale)


class _StockScheduler(BaseMethod):
    """deep_cache.py:17-18 / default_sd.py:15-16 return None from setup_scheduler (stock PNDM); an
    explicit ``scheduler.scheduler_name`` in the YAML overrides that (SURVEY C-12)."""

    def setup_scheduler(self, **kwargs):
        sched = self.config.get("scheduler", None)
        if sched is not None and sched.get("scheduler_name", None):
            return super().setup_scheduler(**kwargs)
        return None


@methods_registry.add_to_registry("deep_cache")
class DeepCacheMethod(_StockScheduler):
    def setup_exp_params(self):
        p = self.config.experiment_params
        self.cache_interval = p.cache_interval
        self.cache_branch_id = p.get("cache_branch_id", 0)
        self.num_inference_steps = p.num_inference_steps

    def run_experiment(self):
        bs = self.config.inference.get("batch_size", 1)
        for interval in self.cache_interval:
            helper = DeepCacheSDHelper(pipe=self.model)
            helper.set_params(cache_interval=interval, cache_branch_id=self.cache_branch_id)
            helper.enable()
            self._new_table()                              # deep_cache.py:39: one table per cache interval
            for steps in self.num_inference_steps:
                self._sweep_point(bs, steps, f"{self.config.experiment_name}, Inference steps: {steps}, "
                                             f"Cache interval: {interval}",
                                  additional_values={"Cache interval": interval})
            helper.disable()


@methods_registry.add_to_registry("default")
class DefaultStableDiffusion(_StockScheduler):
    def setup_exp_params(self):
        self.num_inference_steps = self.config.experiment_params.num_inference_steps

    def run_experiment(self):
        bs = self.config.inference.get("batch_size", 1)
        self._new_table()
        for steps in self.num_inference_steps:
            self._sweep_point(bs, steps, f"{self.config.experiment_name}, Inference steps: {steps}",
                              x0_log_name=f"X0 preds {self.config.experiment_name}, Inference steps: {steps}")


@methods_registry.add_to_registry("two_schedulers")
class TwoSchedulerMethod(BaseMethod):
    def setup_exp_params(self):
        p = self.config.experiment_params
        self.num_inference_steps_first = p.num_inference_steps_first
        self.num_inference_steps_second = p.num_inference_steps_second
        self.num_step_switch = p.num_step_switch
        self.type_switch = p.type_switch
        self.first = {k: p.get(f"first_{k}", "") for k in ("order_solver", "algorithm_type", "final_sigmas_type")}
        self.second = {k: p.get(f"second_{k}", "") for k in ("order_solver", "algorithm_type", "final_sigmas_type")}
        self.batch_size = self.config.inference.get("batch_size", 1)

    def _make(self, name, opts):
        # the reference passes the misspelt `sovler_order` (two_schedulers.py:51,59): dropped by
        # from_config, so the class default order (2) applies; "" = class default (C-3)
        kw = {"sovler_order": opts["order_solver"], "algorithm_type": opts["algorithm_type"],
              "final_sigmas_type": opts["final_sigmas_type"]}
        kw = {k: v for k, v in kw.items() if v != ""}
        return schedulers_registry[name].from_config(self.model.scheduler.config, **kw)

    def setup_scheduler(self):
        self.model.scheduler_first = self._make(self.config.scheduler.scheduler_first, self.first)
        self.model.scheduler_second = self._make(self.config.scheduler.scheduler_second, self.second)

    def run_experiment(self):
        self._new_table()
        for n1, n2, k in zip(self.num_inference_steps_first, self.num_inference_steps_second, self.num_step_switch):
            self._sweep_point(self.batch_size, None,
                              f"{self.config.experiment_name}, Step first: {n1}, Step second: {n2}, Switch: {k}",
                              additional_values={"num_inference_steps_first": n1, "num_inference_steps_second": n2,
                                                 "switch_step": k},
                              x0_grids=True, num_inference_steps_first=n1, num_inference_steps_second=n2,
                              num_step_switch=k, type_switch=self.type_switch)


@methods_registry.add_to_registry("skip_steps")
class SkipStepsMethod(BaseMethod):
    """skip_steps_exp.py:10-144: ``experiment_params.skip_steps`` is a list of lists of LOOP INDICES, zipped with
    ``num_inference_steps`` (one sweep point per pair); the DPM-Solver scheduler is built with
    ``solver_order`` / ``algorithm_type`` / ``final_sigmas_type`` like ``dpm_solver``."""

    def setup_exp_params(self):
        p = self.config.experiment_params
        self.skip_steps = p.skip_steps
        self.num_inference_steps = p.num_inference_steps
        self.solver_order = p.solver_order
        self.algorithm_type = p.algorithm_type
        self.final_sigmas_type = p.final_sigmas_type
        self.batch_size = self.config.inference.get("batch_size", 1)
        if len(self.skip_steps) != len(self.num_inference_steps):
            raise ValueError(f"skip_steps ({len(self.skip_steps)} lists) and num_inference_steps "
                             f"({len(self.num_inference_steps)}) are zipped: they must have the same length")

    def setup_scheduler(self, **kwargs):
        return super().setup_scheduler(solver_order=self.solver_order, algorithm_type=self.algorithm_type,
                                       final_sigmas_type=self.final_sigmas_type)

    def run_experiment(self):
        self._new_table()
        for steps, skip in zip(self.num_inference_steps, self.skip_steps):
            skip = [int(v) for v in skip]
            label = " ".join(map(str, skip))
            name = f"{self.config.experiment_name}, Step main: {steps}, Skip steps:{label}"
            self._sweep_point(self.batch_size, steps, name, x0_log_name=f"X0 preds {name}",
                              additional_values={"num_inference_steps": steps, "skip_steps": label},
                              skip_timesteps=skip)                  # the list itself, skip_steps_exp.py:55-62


@methods_registry.add_to_registry("interliving_schedulers")
class InterlivingSchedulersMethod(BaseMethod):
    """interliving_exp.py:10-188: ``scheduler.scheduler_main`` / ``scheduler.scheduler_inter`` with per-scheduler
    ``{main,inter}_{order_solver,algorithm_type,final_sigmas_type}``; one sweep point per
    (num_inference_steps_first[i], interliving_steps[i]).  "" = class default, as for two_schedulers (C-3)."""

    def setup_exp_params(self):
        p = self.config.experiment_params
        self.num_inference_steps_first = p.num_inference_steps_first
        self.interliving_steps = p.interliving_steps
        self.main = {k: p.get(f"main_{k}", "") for k in ("order_solver", "algorithm_type", "final_sigmas_type")}
        self.inter = {k: p.get(f"inter_{k}", "") for k in ("order_solver", "algorithm_type", "final_sigmas_type")}
        self.batch_size = self.config.inference.get("batch_size", 1)

    def _make(self, name, opts):
        kw = {"solver_order": opts["order_solver"], "algorithm_type": opts["algorithm_type"],
              "final_sigmas_type": opts["final_sigmas_type"]}                # interliving_exp.py:44-47 spells it right
        kw = {k: v for k, v in kw.items() if v != ""}
        return schedulers_registry[name].from_config(self.model.scheduler.config, **kw)

    def setup_scheduler(self):
        self.model.scheduler_main = self._make(self.config.scheduler.scheduler_main, self.main)
        self.model.scheduler_inter = self._make(self.config.scheduler.scheduler_inter, self.inter)

    def run_experiment(self):
        self._new_table()
        for n, inter in zip(self.num_inference_steps_first, self.interliving_steps):
            inter = [int(v) for v in inter]
            self._sweep_point(self.batch_size, n,
                              f"{self.config.experiment_name}, Step main: {n}, Inter steps:{' '.join(map(str, inter))}",
                              additional_values={"num_inference_steps_main": n,
                                                 "num_inter_steps": " ".join(map(str, inter))},
                              x0_grids=True, interliving_steps=inter)
