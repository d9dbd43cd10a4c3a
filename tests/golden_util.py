"""Helpers shared by the golden-fixture tests (tests/golden/*.npz, made by oracle/gen_golden.py
from the unmodified reference)."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def specs_of(fix, key):
    return [tuple(s) for s in ast.literal_eval(str(fix[key]))]


class NumpyLinear:
    """sklearn Pipeline(StandardScaler, LinearRegression).predict from stored parameters."""

    def __init__(self, mean, scale, coef, intercept):
        self.mean, self.scale, self.coef, self.intercept = mean, scale, coef, float(intercept)

    def predict(self, X):
        Xs = (X - self.mean) / self.scale
        return Xs @ self.coef.T + self.intercept


class NumpyForest:
    """Pipeline(StandardScaler, RandomForestRegressor).predict from the packed node arrays."""

    def __init__(self, mean, scale, nodes, bases, root):
        self.mean, self.scale, self.nodes, self.bases, self.root = mean, scale, nodes, bases, root

    def predict(self, X):
        X32 = ((X - self.mean) / self.scale).astype(np.float32)
        n = X.shape[0]
        acc = np.zeros(n)
        thr_all = self.nodes[:, 0].copy().view(np.float32)
        meta_all = self.nodes[:, 1].astype(np.int64)
        val_all = self.nodes.copy().view(np.float64).ravel()
        rows = np.arange(n)
        for t, base in enumerate(self.bases[:len(self.root)]):   # (bases may end with a sentinel)
            idx = np.zeros(n, dtype=np.int64)
            leaf = np.full(n, bool(self.root[t]))
            while not leaf.all():
                live = ~leaf
                g = base + idx[live]
                meta = meta_all[g]
                right = ~(X32[rows[live], meta >> 24] <= thr_all[g])
                idx[live] = (meta & 0x3fffff) + right
                leaf[live] = np.where(right, (meta >> 22) & 1, (meta >> 23) & 1).astype(bool)
            acc += val_all[base + idx]
        return acc / len(self.root)


def regressors_of(fix, kind):
    n_iters = int(fix["n_iters"])
    regs = []
    for i in range(n_iters):
        mean, scale = fix["%s_mean_%d" % (kind, i)], fix["%s_scale_%d" % (kind, i)]
        if kind == "linear":
            regs.append(NumpyLinear(mean, scale, fix["linear_coef_%d" % i],
                                    fix["linear_intercept_%d" % i]))
        else:
            regs.append(NumpyForest(mean, scale, fix["forest_nodes_%d" % i],
                                    fix["forest_bases_%d" % i], fix["forest_root_%d" % i]))
    return regs
