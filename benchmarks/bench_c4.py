#!/usr/bin/env python3
"""Config 4 of BASELINE.json: 3-D voxel gridding of n^3 particles with an NFW-clustered distribution (recipe S2 of
SURVEY 8(d): haloes with NFW profiles + 30 % uniform background, h from the periodic k-NN so it spans decades) onto a
(2n)^3 grid; also the 2-D projection of the same set.  One GPU.  Prints JSON lines."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", dest="n", type=int, default=256)
    ap.add_argument("--haloes", type=int, default=512)
    args = ap.parse_args()
    import torch
    from astro_sph_tools_b200 import synthetic, CoordinateAxes
    from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
    from astro_sph_tools_b200.tools.projections import Gridder3D, Projector2D
    n = args.n; N = n ** 3
    t0 = time.time()
    pos, _ = synthetic.s2_positions(N, 1.0, n_haloes=args.haloes, seed=12345)
    print(json.dumps({"generated": N, "seconds": round(time.time() - t0, 1)}), flush=True)
    pos_d = torch.from_numpy(pos).cuda()
    sol = SmoothingLengthSolver()
    sol.solve(pos_d[:100000].contiguous(), 48, 1.0)          # loads the kernels (cold start is not part of the number)
    sol.solve(pos_d, 48, 1.0)
    torch.cuda.synchronize(); t0 = time.time()
    h_d = sol.solve(pos_d, 48, 1.0)
    torch.cuda.synchronize(); t_knn = time.time() - t0
    h = h_d.cpu().numpy()
    print(json.dumps({"config": f"k-NN k=48 periodic on NFW-clustered {n}^3", "seconds": round(t_knn, 3), "queries_per_s": N / t_knn,
                      "h_min": float(h.min()), "h_median": float(np.median(h)), "h_max": float(h.max())}), flush=True)
    from scipy.spatial import cKDTree
    sel = np.random.default_rng(1).choice(N, 5000, replace=False)
    ref = cKDTree(pos, boxsize=1.0).query(pos[sel], k=48, workers=-1)[0][:, 47]
    print(json.dumps({"knn_bit_equal_to_scipy_on_5000_queries": bool(np.array_equal(h[sel], ref))}), flush=True)
    sol._ws = None; torch.cuda.empty_cache()
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    g = Gridder3D()
    size = (2 * n,) * 3
    out = torch.empty(size, dtype=torch.float64, device="cuda")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        g.grid(pos_d, h_d, m_d, size, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out, timing=(it == 1))
        torch.cuda.synchronize(); dt = time.time() - t0
    print(json.dumps({"config": f"C4: NFW-clustered {n}^3 -> {2 * n}^3 voxels, periodic, h = d_48", "seconds": round(dt, 4),
                      "particles_per_s": N / dt, "stats": g.last_stats, "mass_sum": float(out.sum().item()) / (2 * n) ** 3}), flush=True)
    del out; g._ws = None; torch.cuda.empty_cache()
    eng = Projector2D()
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        img = eng.project(pos_d, h_d, m_d, (8 * n, 8 * n), CoordinateAxes.Z, (0, 1, 0, 1), "cubic_spline_3d", True, 1.0, timing=(it == 1))
        torch.cuda.synchronize(); dt = time.time() - t0
    print(json.dumps({"config": f"2-D projection of the same set -> {8 * n}^2, periodic", "seconds": round(dt, 4), "particles_per_s": N / dt,
                      "stats": eng.last_stats}), flush=True)


if __name__ == "__main__":
    main()
