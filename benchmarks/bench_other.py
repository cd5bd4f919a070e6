#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json on one GPU (not the driver's bench line; see bench.py for that):
  C5 slice: smoothing-length k-NN (k = 48, periodic) on n^3 particles   -> queries/s, parity vs scipy on a sample
  C4:       3-D voxel gridding of n^3 particles onto a (2n)^3 grid       -> particles/s
  C1:       64^3 -> 512^2 Wendland-C2 periodic surface density           -> particles/s
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
    ap.add_argument("--which", default="knn,grid,c1")
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


if __name__ == "__main__":
    main()
