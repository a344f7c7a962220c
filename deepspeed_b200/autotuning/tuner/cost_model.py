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
