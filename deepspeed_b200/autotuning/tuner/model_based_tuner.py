"""Model-based tuner (reference ``tuner/model_based_tuner.py``): evaluate a few random configs, fit a cost model on
the flattened numeric features, then repeatedly run the config the model ranks best (with epsilon exploration)."""
import numbers
import random

from ..utils import flatten
from .base_tuner import BaseTuner
from .cost_model import XGBoostCostModel

INIT_NUM = 2


class ModelBasedTuner(BaseTuner):

    def __init__(self, exps, resource_manager, metric, tuning_space=None):
        super().__init__(exps, resource_manager, metric)
        # boosted trees (the reference's XGBoost model, in tree) once enough configurations were measured; ridge before
        self.cost_model = XGBoostCostModel("reg")
        self.visited, self.evaluated = set(), []
        self.keys = sorted({k for e in exps for k, v in flatten(e["ds_config"]).items()
                            if isinstance(v, (numbers.Number, bool))})
        self.random_exploration_ratio = 0.2
        self._trained = False

    def _feat(self, exp):
        f = flatten(exp["ds_config"])
        return [float(f.get(k, 0) or 0) for k in self.keys]

    def next_batch(self, sample_size=1):
        out = []
        for _ in range(min(sample_size, len(self.all_exps))):
            if len(self.evaluated) < INIT_NUM or not self._trained or random.random() < self.random_exploration_ratio:
                idx = random.randrange(len(self.all_exps))
            else:
                preds = self.cost_model.predict([self._feat(e) for e in self.all_exps])
                idx = int(preds.argmin() if self.metric == "latency" else preds.argmax())
            e = self.all_exps.pop(idx)
            out.append(e)
            self._pending = getattr(self, "_pending", []) + [e]
        return out

    def update(self):
        for e in getattr(self, "_pending", []):
            if e.get("result") is not None:
                self.evaluated.append((self._feat(e), e["result"]))
        self._pending = []
        if len(self.evaluated) >= INIT_NUM and self.keys:
            xs, ys = zip(*self.evaluated)
            self.cost_model.fit(xs, ys)
            self._trained = True
