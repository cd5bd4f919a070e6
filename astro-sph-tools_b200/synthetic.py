"""Seeded synthetic particle sets used by tests and bench.py.

S1 "jittered lattice": rng = default_rng(seed); lattice (i+0.5)/n*L in C order; pos = mod(lattice +
normal(0, 0.2 L/n), L); masses 1/N; T = 10**uniform(4,7).  h is the reference convention
(io/SWIFT/_SnapshotSWIFT.py:62-83): distance to the k-th neighbour, self included.
S2 "NFW-clustered": mixture of haloes with NFW radial profile + 30 % uniform background.

These are input generators only (numpy on the host); they are not part of the timed path.
"""
import numpy as np


def s1_positions(n, L=1.0, seed=12345):
    rng = np.random.default_rng(seed)
    g = (np.arange(n, dtype=np.float64) + 0.5) / n * L
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = np.mod(lattice + rng.normal(0.0, 0.2 * L / n, lattice.shape), L)
    pos[pos >= L] = 0.0                      # mod can return L for tiny negative inputs
    return np.ascontiguousarray(pos), rng


def s1_blocks(n, L=1.0, seed=12345, planes=16):
    """The same particle set as s1_positions, produced as consecutive row blocks (i0, i1, pos[i0:i1]) of `planes` lattice
    x-planes each: Generator.normal fills its output in C order from one sequential stream, so drawing the jitter block by
    block yields the identical numbers with bounded memory (512^3: 3.2 GB whole, 100 MB per 16-plane block)."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n, dtype=np.float64) + 0.5) / n * L
    for p0 in range(0, n, planes):
        p1 = min(n, p0 + planes)
        lattice = np.stack(np.meshgrid(g[p0:p1], g, g, indexing="ij"), axis=-1).reshape(-1, 3)
        blk = np.mod(lattice + rng.normal(0.0, 0.2 * L / n, lattice.shape), L)
        blk[blk >= L] = 0.0
        yield p0 * n * n, p1 * n * n, blk


def s1_h_lattice_estimate(n, k=48, L=1.0):
    """Mean d_k of a Poisson-like set of number density n^3/L^3: radius of the sphere holding k points."""
    return (3.0 * k / (4.0 * np.pi)) ** (1.0 / 3.0) * L / n


def s1(n, k=48, L=1.0, seed=12345, h_mode="scipy", h_scale=1.0, with_temperature=False, knn=None):
    """Returns dict(pos (N,3), h (N,), mass (N,), [T (N,)]).

    h_mode: 'scipy'   -> scipy.spatial.cKDTree(boxsize=L).query (host, slow for n >= 128)
            'callable'-> knn(pos, k, L) supplied by the caller (e.g. the CUDA k-NN of this package)
            'uniform' -> the constant lattice estimate (no neighbour search)
    """
    pos, rng = s1_positions(n, L, seed)
    N = pos.shape[0]
    if h_mode == "scipy":
        from scipy.spatial import cKDTree
        h = cKDTree(pos, boxsize=L).query(pos, k=k, workers=-1)[0][:, k - 1].copy()
    elif h_mode == "callable":
        h = np.asarray(knn(pos, k, L), dtype=np.float64)
    elif h_mode == "uniform":
        h = np.full(N, s1_h_lattice_estimate(n, k, L))
    else:
        raise ValueError(h_mode)
    out = dict(pos=pos, h=h * h_scale, mass=np.full(N, 1.0 / N), L=L)
    if with_temperature:
        out["T"] = 10.0 ** rng.uniform(4.0, 7.0, N)
    return out


def _nfw_radius(u, c):
    """Inverse CDF of the NFW enclosed mass m(x) = ln(1+x) - x/(1+x), x = r/rs in [0, c], by bisection."""
    m = lambda x: np.log1p(x) - x / (1.0 + x)
    target = u * m(c)
    lo = np.zeros_like(u); hi = np.full_like(u, c)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        big = m(mid) > target
        hi = np.where(big, mid, hi); lo = np.where(big, lo, mid)
    return 0.5 * (lo + hi)


def s2_positions(N, L=1.0, n_haloes=64, seed=12345, background=0.3):
    rng = np.random.default_rng(seed)
    n_bg = int(N * background)
    n_h = N - n_bg
    centres = rng.uniform(0, L, (n_haloes, 3))
    conc = rng.uniform(5.0, 10.0, n_haloes)
    rvir = L * 0.04 * rng.uniform(0.5, 1.5, n_haloes)
    which = rng.integers(0, n_haloes, n_h)
    x = _nfw_radius(rng.uniform(0, 1, n_h), conc[which])
    r = x / conc[which] * rvir[which]
    v = rng.normal(size=(n_h, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    pos = np.concatenate([centres[which] + r[:, None] * v, rng.uniform(0, L, (n_bg, 3))])
    pos = np.mod(pos, L); pos[pos >= L] = 0.0
    return np.ascontiguousarray(pos[rng.permutation(N)]), rng


def s2(N, k=48, L=1.0, n_haloes=64, seed=12345, h_mode="scipy", knn=None, h_max=None):
    pos, rng = s2_positions(N, L, n_haloes, seed)
    if h_mode == "scipy":
        from scipy.spatial import cKDTree
        h = cKDTree(pos, boxsize=L).query(pos, k=k, workers=-1)[0][:, k - 1].copy()
    else:
        h = np.asarray(knn(pos, k, L), dtype=np.float64)
    if h_max is not None:
        h = np.minimum(h, h_max)
    return dict(pos=pos, h=h, mass=np.full(N, 1.0 / N), L=L)


def add_periodic_ghosts(pos, h, props, L, cols=(0, 1)):
    """Reference-side recipe for a periodic projection: replicate every particle
    whose kernel support (2h) crosses a face of the box in one of the in-plane columns."""
    P = [pos]; H = [h]; A = [props]
    a, b = cols
    for ia in (-1, 0, 1):
        for ib in (-1, 0, 1):
            if ia == 0 and ib == 0:
                continue
            sel = np.ones(len(h), dtype=bool)
            if ia == -1: sel &= pos[:, a] + 2 * h > L
            if ia == 1: sel &= pos[:, a] - 2 * h < 0
            if ib == -1: sel &= pos[:, b] + 2 * h > L
            if ib == 1: sel &= pos[:, b] - 2 * h < 0
            q = pos[sel].copy(); q[:, a] += ia * L; q[:, b] += ib * L
            P.append(q); H.append(h[sel]); A.append(props[..., sel])
    return np.concatenate(P), np.concatenate(H), np.concatenate(A, axis=-1)
