"""Synthetic ionisation-like tables and gas states shared by the table tests (HM01-like shape: log10 nH x log10 T x z,
io/ionisation_tables/_HM01.py:73-97; values = log10 ion fraction <= 0)."""
import numpy as np


def synthetic_table(seed=7, shape=(41, 141, 49), uniform=True):
    rng = np.random.default_rng(seed)
    if uniform:
        axes = [np.linspace(-8.0, 2.0, shape[0]), np.linspace(2.0, 9.0, shape[1]), np.linspace(0.0, 8.989, shape[2])]
    else:
        axes = [np.sort(rng.uniform(-8, 2, shape[0])), np.sort(rng.uniform(2, 9, shape[1])), np.sort(rng.uniform(0, 9, shape[2]))]
    a, b, c = np.meshgrid(*axes, indexing="ij")
    table = -np.abs(0.3 * (a + 3) ** 2 + 1.5 * np.sin(b) * (b - 5.5) + 0.2 * c) + 0.01 * rng.normal(size=shape)
    return np.minimum(table, 0.0), axes


def gas_state(seed, n, axes, with_edge_cases=True):
    rng = np.random.default_rng(seed)
    lo = np.array([a[0] for a in axes]); hi = np.array([a[-1] for a in axes])
    x = rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo), (n, len(axes)))     # ~14 % of the points fall outside
    if with_edge_cases and n >= 64:
        for d, a in enumerate(axes):                     # exact grid points (incl. both ends) in every dimension
            x[8 * d:8 * d + 8, d] = a[[0, -1, 1, -2, len(a) // 2, 3, 0, -1]]
        nd = len(axes)
        x[40, 0] = np.nan
        x[41, nd - 1] = np.nan; x[41, 0] = lo[0] - 1.0 if nd > 1 else np.nan       # NaN wins over out-of-bounds
        x[42] = lo; x[43] = hi
        x[44, nd - 1] = np.inf; x[45, nd - 1] = -np.inf
    return x
