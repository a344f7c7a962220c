"""Tuner loop (reference ``autotuning/tuner/base_tuner.py``): ask for a batch of experiments, run them through the
resource manager, track the best, stop early after ``early_stopping`` non-improving trials."""
import sys

from deepspeed_b200.utils.logging import logger


class BaseTuner:

    def __init__(self, exps, resource_manager, metric):
        self.all_exps = exps
        self.rm = resource_manager
        self.best_iter = 0
        self.best_exp = None
        self.best_metric_val = None
        self.metric = metric or "throughput"
        logger.info(f"total number of exps =  {len(self.all_exps)}")

    def has_next(self):
        return len(self.all_exps) > 0

    def next_batch(self, sample_size):
        raise NotImplementedError

    def update(self):
        pass

    def _better(self, a, b):
        if b is None:
            return True
        return a < b if self.metric == "latency" else a > b

    def tune(self, sample_size=1, n_trials=1000, early_stopping=None):
        i = 0
        try:
            while i < n_trials and self.has_next():
                batch = self.next_batch(sample_size)
                paths = self.rm.schedule_experiments_dicts(batch)
                self.rm.run()
                for exp in batch:
                    val = self.rm.metric_of(exp, self.metric)
                    exp["result"] = val
                    if val is not None and self._better(val, self.best_metric_val):
                        self.best_exp, self.best_metric_val, self.best_iter = exp, val, i
                i += len(batch)
                self.update()
                self.rm.clear()
                if early_stopping and i >= self.best_iter + early_stopping:
                    logger.info(f"Tuner early stopped at iteration {i}. Best iteration is {self.best_iter}.")
                    break
            return i
        except Exception:
            logger.info(f"Tuner error: {sys.exc_info()[0]}")
            return i
