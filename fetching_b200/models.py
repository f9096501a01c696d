"""Synthetic data sets of the BASELINE.json shapes and theta helpers.

Pure numpy input generation -- no likelihood arithmetic lives here.  The recipes follow the
reference where it has one:
  C1  normal    samplers/normal.cc:41-46 (default-seeded mt19937 + Box-Muller) via `normal_dataset`
  C3/C5 mixture pysamplers/pymixture.py:54-62 with d=2: c ~ U{0..K-1}, x ~ N(pos_c, I), pos = 4 U - 2
  C4  lasso     X ~ N(0,1) column-standardised, y = X beta* + N(0,1), centred (pyblasso.py:60-61)
  C2  Boltzmann W dense symmetric N(0, 1/D), X ~ N(0,1) (new workload)
"""
from __future__ import annotations

import numpy as np


class MT19937:
    """boost::random::mt19937 (== std::mt19937), default seed 5489: the raw 32-bit stream the
    reference's RandomSequence replays (randomsequence.hh:19).  numpy's legacy generator is the
    same twister; only the seeding call differs (init_genrand)."""

    def __init__(self, seed: int = 5489):
        self._rs = np.random.RandomState()
        mt = np.empty(624, dtype=np.uint32)
        mt[0] = seed
        for i in range(1, 624):
            mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
        self._rs.set_state(("MT19937", mt, 624))

    def raw(self, n: int) -> np.ndarray:
        return self._rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)


def normal_dataset(n_points: int) -> np.ndarray:
    """samplers/normal.cc:41-46: value i = N(0,1) + (i odd ? 100 : 200); N(0,1) by Box-Muller with
    a cached second variate over uniform_01 = raw * 2^-32 (Boost <= 1.63, see oracle/shim)."""
    n = 2 * n_points
    pairs = (n + 1) // 2
    u = MT19937().raw(2 * pairs).astype(np.float64) * (1.0 / 4294967296.0)
    r1, r2 = u[0::2], u[1::2]
    rho = np.sqrt(-2.0 * np.log(1.0 - r2))
    ang = 2.0 * 3.14159265358979323846 * r1
    z = np.empty(2 * pairs)
    z[0::2] = rho * np.cos(ang)
    z[1::2] = rho * np.sin(ang)
    z = z[:n] * 1.0 + 0.0
    z[0::2] += 200.0
    z[1::2] += 100.0
    return z.reshape(n_points, 2)


def mixture_dataset(n_rows: int, K: int = 8, d: int = 2, seed: int = 0, length: float = 4.0):
    """(data[n][d], positions[K][d]) -- the pymixture.py:54-62 recipe at a chosen d."""
    rng = np.random.RandomState(seed)
    pos = length * rng.random_sample((K, d)) - length / 2.0
    comp = (rng.random_sample(n_rows) * K).astype(np.int64)
    data = rng.standard_normal((n_rows, d)) + pos[comp]
    return data, pos


def mixture_thetas(positions: np.ndarray, B: int, seed: int = 1, scale: float = 0.05) -> np.ndarray:
    """B frontier nodes: label-permuted, jittered component means (pymixture.py:81-84 style)."""
    rng = np.random.RandomState(seed)
    K, d = positions.shape
    out = np.empty((B, K * d))
    for b in range(B):
        out[b] = (positions[rng.permutation(K)] + rng.normal(0, scale, (K, d))).ravel()
    return out


def blasso_dataset(n_rows: int, p: int, seed: int = 0, nnz: int = 20):
    """(X standardised, y centred, beta*)."""
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n_rows, p))
    X = (X - X.mean(axis=0)) / X.std(axis=0)
    beta = np.zeros(p)
    idx = rng.choice(p, size=min(nnz, p), replace=False)
    beta[idx] = rng.normal(0, 1.0, idx.size)
    y = X @ beta + rng.standard_normal(n_rows)
    y = y - y.mean()
    return X, y, beta


def blasso_thetas(beta: np.ndarray, B: int, seed: int = 1, scale: float = 0.01) -> np.ndarray:
    rng = np.random.RandomState(seed)
    out = np.empty((B, beta.size + 1))
    for b in range(B):
        out[b, 0] = abs(1.0 + rng.normal(0, 0.05))
        out[b, 1:] = beta + rng.normal(0, scale, beta.size)
    return out


def boltzmann_dataset(D: int, B: int, seed: int = 1):
    rng = np.random.RandomState(seed)
    W = rng.normal(0, 1.0 / np.sqrt(D), (D, D))
    W = (W + W.T) / 2.0
    X = rng.standard_normal((B, D))
    return W, X


def normal_thetas(B: int, seed: int = 1, center=(0.0, 0.0), scale: float = 0.1) -> np.ndarray:
    """B nodes around `center` with identity covariance: what normalsampler.hh:66-79 proposes."""
    rng = np.random.RandomState(seed)
    out = np.zeros((B, 5))
    out[:, 0] = center[0] + rng.normal(0, scale, B)
    out[:, 1] = center[1] + rng.normal(0, scale, B)
    out[:, 2] = 1.0
    out[:, 4] = 1.0
    return out
