"""Cost model of the model-based tuner.  The reference fits XGBoost (``tuner/cost_model.py``); here a closed-form
ridge regression on (normalised features + pairwise products) — no external dependency, good enough to rank a few
dozen configurations."""
import numpy as np


class RidgeCostModel:

    def __init__(self, l2=1e-2):
        self.l2 = l2
        self.w = None
        self.mu = self.sd = None

    def _expand(self, X):
        X = (np.asarray(X, dtype=np.float64) - self.mu) / self.sd
        n, d = X.shape
        cross = [X[:, i:i + 1] * X[:, j:j + 1] for i in range(d) for j in range(i, d)]
        return np.concatenate([np.ones((n, 1)), X] + cross, axis=1)

    def fit(self, xs, ys):
        X = np.asarray(xs, dtype=np.float64)
        self.mu, self.sd = X.mean(0), X.std(0) + 1e-9
        y = np.asarray(ys, dtype=np.float64)
        self.scale = max(np.abs(y).max(), 1e-9)
        F = self._expand(X)
        A = F.T @ F + self.l2 * np.eye(F.shape[1])
        self.w = np.linalg.solve(A, F.T @ (y / self.scale))

    def predict(self, xs):
        return self._expand(np.asarray(xs, dtype=np.float64)) @ self.w * self.scale


class XGBoostCostModel:
    """Reference-named cost model (``tuner/cost_model.py``): gradient-boosted ranking when ``xgboost`` is installed, the
    closed-form ridge model otherwise -- same ``fit`` / ``predict`` contract."""

    def __init__(self, loss_type="reg", num_threads=None, log_interval=25, upper_model=None):
        assert loss_type in ("reg", "rank")
        self.loss_type, self.num_threads = loss_type, num_threads
        try:
            import xgboost  # noqa: F401
            self._xgb = xgboost
        except ImportError:
            self._xgb = None
        self._fallback = RidgeCostModel()
        self.bst = None

    def fit(self, xs, ys):
        if self._xgb is None:
            self._fallback.fit(xs, ys)
            return
        x, y = np.asarray(xs, dtype=np.float32), np.asarray(ys, dtype=np.float32)
        y = y / max(float(np.max(y)), 1e-9)
        params = {"max_depth": 3, "gamma": 1e-4, "min_child_weight": 1, "subsample": 1.0, "eta": 0.3, "lambda": 1.0,
                  "alpha": 0, "objective": "reg:linear" if self.loss_type == "reg" else "rank:pairwise", "verbosity": 0}
        if self.num_threads:
            params["nthread"] = self.num_threads
        self.bst = self._xgb.train(params, self._xgb.DMatrix(x, y), num_boost_round=10)

    def predict(self, xs):
        if self._xgb is None or self.bst is None:
            return self._fallback.predict(xs)
        return self.bst.predict(self._xgb.DMatrix(np.asarray(xs, dtype=np.float32)))
