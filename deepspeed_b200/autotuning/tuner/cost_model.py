"""Cost models of the model-based tuner.  The reference fits XGBoost (``tuner/cost_model.py``: depth-3 trees, 10 boosting
rounds, regression or pairwise-rank objective).  ``BoostedTreesCostModel`` is that model written out in numpy (exact greedy
split search, shrinkage, L2-regularised leaf values -- there are a few dozen samples, not millions), so the tuner behaves
the same with no external dependency; ``RidgeCostModel`` (normalised features + pairwise products, closed form) is the
fallback for fewer than a handful of measured points.  ``XGBoostCostModel`` keeps the reference's name and uses the real
library when it happens to be installed."""
import numpy as np


class _Tree:
    """One regression tree on gradients / hessians (second-order boosting, like XGBoost's exact method)."""

    def __init__(self, max_depth, lam, gamma, min_child_weight):
        self.max_depth, self.lam, self.gamma, self.mcw = max_depth, lam, gamma, min_child_weight
        self.nodes = []  # (feature, threshold, left, right) or (None, value, None, None)

    def fit(self, X, g, h):
        self._grow(X, g, h, np.arange(len(g)), 0)
        return self

    def _leaf(self, g, h):
        self.nodes.append((None, -g.sum() / (h.sum() + self.lam), None, None))
        return len(self.nodes) - 1

    def _grow(self, X, g, h, idx, depth):
        G, H = g[idx].sum(), h[idx].sum()
        if depth >= self.max_depth or len(idx) < 2:
            return self._leaf(g[idx], h[idx])
        best = (self.gamma, None, None)
        parent = G * G / (H + self.lam)
        for f in range(X.shape[1]):
            order = idx[np.argsort(X[idx, f], kind="stable")]
            xs = X[order, f]
            gl, hl = np.cumsum(g[order])[:-1], np.cumsum(h[order])[:-1]
            ok = (xs[1:] > xs[:-1]) & (hl >= self.mcw) & (H - hl >= self.mcw)
            if not ok.any():
                continue
            gain = 0.5 * (gl * gl / (hl + self.lam) + (G - gl)**2 / (H - hl + self.lam) - parent)
            gain[~ok] = -np.inf
            k = int(np.argmax(gain))
            if gain[k] > best[0]:
                best = (gain[k], f, 0.5 * (xs[k] + xs[k + 1]))
        if best[1] is None:
            return self._leaf(g[idx], h[idx])
        _, f, thr = best
        me = len(self.nodes)
        self.nodes.append(None)
        left = self._grow(X, g, h, idx[X[idx, f] <= thr], depth + 1)
        right = self._grow(X, g, h, idx[X[idx, f] > thr], depth + 1)
        self.nodes[me] = (f, thr, left, right)
        return me

    def predict(self, X):
        out = np.empty(len(X))
        for i, x in enumerate(X):
            n = 0
            while self.nodes[n][0] is not None:
                f, thr, l, r = self.nodes[n]
                n = l if x[f] <= thr else r
            out[i] = self.nodes[n][1]
        return out


class BoostedTreesCostModel:
    """Gradient-boosted regression trees with XGBoost's defaults from the reference tuner: ``max_depth=3``, ``eta=0.3``,
    ``lambda=1``, ``gamma=1e-4``, 10 rounds.  ``loss_type="rank"`` optimises the pairwise logistic ranking loss (what matters
    to the tuner is the ORDER of configurations), ``"reg"`` squared error on the max-normalised metric."""

    def __init__(self, loss_type="reg", rounds=10, max_depth=3, eta=0.3, lam=1.0, gamma=1e-4, min_child_weight=1.0):
        assert loss_type in ("reg", "rank")
        self.loss_type, self.rounds, self.eta = loss_type, rounds, eta
        self.tree_args = (max_depth, lam, gamma, min_child_weight)
        self.trees, self.base = [], 0.0

    def _grad(self, pred, y):
        if self.loss_type == "reg":
            return pred - y, np.ones_like(y)
        g, h = np.zeros_like(y), np.zeros_like(y)
        for i in range(len(y)):  # pairwise logistic: every (better, worse) pair pushes the two scores apart
            for j in range(len(y)):
                if y[i] > y[j]:
                    p = 1.0 / (1.0 + np.exp(pred[i] - pred[j]))
                    g[i] -= p
                    g[j] += p
                    h[i] += p * (1 - p)
                    h[j] += p * (1 - p)
        return g, np.maximum(h, 1e-6)

    def fit(self, xs, ys):
        X = np.asarray(xs, dtype=np.float64)
        y = np.asarray(ys, dtype=np.float64)
        y = y / max(float(np.max(np.abs(y))), 1e-9)
        self.base = float(y.mean()) if self.loss_type == "reg" else 0.0
        self.trees = []
        pred = np.full(len(y), self.base)
        for _ in range(self.rounds):
            g, h = self._grad(pred, y)
            t = _Tree(*self.tree_args).fit(X, g, h)
            self.trees.append(t)
            pred = pred + self.eta * t.predict(X)

    def predict(self, xs):
        X = np.asarray(xs, dtype=np.float64)
        out = np.full(len(X), self.base)
        for t in self.trees:
            out = out + self.eta * t.predict(X)
        return out


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
    """Reference-named cost model (``tuner/cost_model.py``): the ``xgboost`` library when it is installed, the in-tree
    boosted-trees model with the same hyper-parameters otherwise (ridge below 6 samples) -- same ``fit`` / ``predict``
    contract."""

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
            self._fallback = BoostedTreesCostModel(self.loss_type) if len(ys) >= 6 else RidgeCostModel()
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
