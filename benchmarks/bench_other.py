#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json on one GPU (not the driver's bench line; see bench.py for that):
  C5 slice: smoothing-length k-NN (k = 48, periodic) on n^3 particles   -> queries/s, parity vs scipy on a sample
  C4:       3-D voxel gridding of n^3 particles onto a (2n)^3 grid       -> particles/s
  C1:       64^3 -> 512^2 Wendland-C2 periodic surface density           -> particles/s
  ion:      HM01-shaped table lookup fused into ion weights (8(f) N3)    -> particles/s, GB/s; scipy timed beside it
  halo:     nearest-halo lookup (8(f) N2): periodic 1-NN of the n^3 particles among 10^6 halo centres -> queries/s; scipy beside it
Prints one JSON object per configuration."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=3, warm=1):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--which", default="knn,grid,c1,ion,halo")
    args = ap.parse_args()
    import torch
    from astro_sph_tools_b200 import synthetic, CoordinateAxes
    from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
    from astro_sph_tools_b200.tools.projections import Gridder3D, Projector2D
    peak = 6554.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n = args.n
    pos, rng = synthetic.s1_positions(n)
    N = pos.shape[0]
    pos_d = torch.from_numpy(pos).cuda()
    sol = SmoothingLengthSolver()
    h_d = None
    if "knn" in args.which:
        ms = timed(lambda: sol.solve(pos_d, 48, 1.0), reps=2)
        h_d = sol.solve(pos_d, 48, 1.0)
        # parity on a sample of queries against scipy (the reference's arithmetic)
        from scipy.spatial import cKDTree
        sel = np.random.default_rng(1).choice(N, 20000, replace=False)
        ref = cKDTree(pos, boxsize=1.0).query(pos[sel], k=48, workers=-1)[0][:, 47]
        ok = bool(np.array_equal(h_d.cpu().numpy()[sel], ref))
        print(json.dumps({"config": f"k-NN k=48 periodic, S1 {n}^3 = {N} particles, 1 GPU", "ms": ms, "queries_per_s": N / (ms * 1e-3),
                          "algorithmic_bytes": N * 32, "hbm_frac": N * 32 / (ms * 1e-3) / 1e9 / peak,
                          "bit_equal_to_scipy_on_20000_queries": ok}))
    if h_d is None:
        h_d = torch.full((N,), synthetic.s1_h_lattice_estimate(n, 48), dtype=torch.float64, device="cuda")
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    if "grid" in args.which:
        g = Gridder3D()
        size = (2 * n,) * 3
        out = torch.empty(size, dtype=torch.float64, device="cuda")
        ms = timed(lambda: g.grid(pos_d, h_d, m_d, size, (0, 0, 0), (1, 1, 1), out=out), reps=2)
        g.grid(pos_d, h_d, m_d, size, (0, 0, 0), (1, 1, 1), out=out, timing=True)
        alg = N * 40 + (2 * n) ** 3 * 8
        print(json.dumps({"config": f"3-D grid S1 {n}^3 particles -> {2 * n}^3 voxels, h = d_48, cubic spline", "ms": ms,
                          "particles_per_s": N / (ms * 1e-3), "algorithmic_bytes": alg, "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                          "pairs": g.last_stats["n_pairs"], "rounds": g.last_stats["n_rounds"], "stage_ms": g.last_stats["stage_ms"],
                          "mass_sum": float(out.sum().item() / (2 * n) ** 3)}))
    if "c1" in args.which:
        s = synthetic.s1(64, k=48, h_mode="callable", knn=lambda p, k, L: sol.solve(torch.from_numpy(p).cuda(), k, L).cpu().numpy())
        eng = Projector2D()
        p_d, hh, mm = (torch.from_numpy(s[k]).cuda() for k in ("pos", "h", "mass"))
        f = lambda: eng.project(p_d, hh, mm, (512, 512), CoordinateAxes.Z, (0, 1, 0, 1), "wendland_c2_2d", True, 1.0)
        ms = timed(f, reps=5)
        img = f()
        print(json.dumps({"config": "C1: S1 64^3 -> 512^2 Wendland C2 periodic surface density", "ms": ms, "particles_per_s": 64 ** 3 / (ms * 1e-3),
                          "mass_conservation_rel": abs(float(img.sum().item()) / 512 ** 2 - 1.0)}))
    if "halo" in args.which:
        # SURVEY 8(f) N2 (_scripts/find_nearest_haloes.py:207-215): KDTree(centres, boxsize=L).query(particle positions)
        from scipy.spatial import cKDTree
        n_halo = 1_000_000
        centres = np.random.default_rng(9).uniform(0.0, 1.0, (n_halo, 3))
        c_d = torch.from_numpy(centres).cuda()
        f = lambda: sol.query(c_d, pos_d, 1, 1.0)
        ms = timed(f, reps=3)
        dist, idx = f()
        sel = np.random.default_rng(2).choice(N, 200000, replace=False)
        tree = cKDTree(centres, boxsize=1.0)
        t0 = time.perf_counter(); d_ref, i_ref = tree.query(pos[sel], k=1, workers=1); t1 = time.perf_counter() - t0
        t0 = time.perf_counter(); tree.query(pos[sel], k=1, workers=-1); ta = time.perf_counter() - t0
        ok = bool(np.array_equal(dist.cpu().numpy()[sel, 0], d_ref) and np.array_equal(idx.cpu().numpy()[sel, 0], i_ref))
        print(json.dumps({"config": f"nearest-halo lookup: periodic 1-NN of S1 {n}^3 = {N} particles among {n_halo} uniformly placed centres",
                          "ms": ms, "queries_per_s": N / (ms * 1e-3), "bit_equal_to_scipy_on_sample": ok,
                          "scipy_queries_per_s_1_core": len(sel) / t1, "scipy_queries_per_s_all_cores": len(sel) / ta, "cores": os.cpu_count(),
                          "roofline": {"bound": "hbm", "algorithmic_bytes": N * (24 + 8 + 4) + n_halo * 24,
                                       "frac": (N * 36 + n_halo * 24) / (ms * 1e-3) / 1e9 / peak}}), flush=True)
        del c_d, dist, idx
    if "ion" in args.which:
        from astro_sph_tools_b200.tools.ionisation import IonisationTableBase
        r = np.random.default_rng(3)
        axes = [np.linspace(-8.0, 2.0, 41), np.linspace(2.0, 9.0, 141), np.linspace(0.0, 8.989, 49)]       # HM01 shape
        table = -np.abs(r.normal(size=(41, 141, 49))) * 3
        t = IonisationTableBase(table, *axes, redshift_input_index=2)
        lognh = torch.from_numpy(r.uniform(-8.2, 2.1, N)).cuda(); logt = torch.from_numpy(r.uniform(1.9, 9.1, N)).cuda()
        out = torch.empty(N, dtype=torch.float64, device="cuda")
        ms = timed(lambda: t.device_eval([lognh, logt, None], redshift=2.2, base=m_d, pow10=True, out=out), reps=10, warm=3)
        alg = N * 32                                                         # log nH, log T, element mass in; weight out
        ns = min(N, 2_000_000)
        x = np.stack([lognh[:ns].cpu().numpy(), logt[:ns].cpu().numpy(), np.full(ns, 2.2)], axis=1)
        from scipy.interpolate import RegularGridInterpolator
        rgi = RegularGridInterpolator(tuple(axes), table, bounds_error=False, fill_value=-np.inf)
        t0 = time.perf_counter(); ref = rgi(x); t_cpu = time.perf_counter() - t0
        ok = bool(np.array_equal(t.device_eval([lognh[:ns], logt[:ns], None], redshift=2.2).cpu().numpy(), ref, equal_nan=True))
        print(json.dumps({"config": f"ion weights: 41x141x49 table (HM01 shape), {N} particles, mass * 10**table(lognH, logT, z)",
                          "ms": ms, "particles_per_s": N / (ms * 1e-3), "algorithmic_bytes": alg,
                          "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak, "bit_equal_to_scipy": ok,
                          "scipy_particles_per_s_1_core": ns / t_cpu}))


if __name__ == "__main__":
    main()
