#!/usr/bin/env python3
"""Small run of every entry point (run it with AST_DEBUG_SYNC=1 so that every kernel is synchronised and named on failure;
compute-sanitizer is closed on this pool): 2-D projection (direct, tiled, large-h list, periodic,
several rounds, batched host path), 3-D grid, k-NN (both kernels, subsets with the reach-limited build, lists), ion table."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200.tools.projections import Projector2D, Gridder3D
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.ionisation import IonisationTableBase

rng = np.random.default_rng(0)
n = 6000
pos = rng.uniform(0, 1, (n, 3)); h = np.exp(rng.uniform(np.log(0.0005), np.log(0.2), n)); m = rng.uniform(0.5, 1.5, n)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for kw in ({}, dict(small_max_px=1, huge_min_tiles=0), dict(pair_capacity=5000)):
    eng = Projector2D(**kw)
    for per in (False, True):
        out = eng.project(d(pos), d(h), [d(m), d(2 * m)], (150, 130), 2, (0.0, 1.0, 0.1, 0.9), "cubic_spline_3d", per, 1.0 if per else None)
        assert torch.isfinite(out).all()
eng = Projector2D()
img = eng.project_host(pos, h, m, (150, 130), 1, (0.0, 1.0, 0.0, 1.0), "wendland_c2_2d", batch_particles=1000)
g = Gridder3D(pair_capacity=20000)
vol = g.grid(d(pos), d(np.minimum(h, 0.08)), d(m), (40, 36, 44), (0, 0, 0), (1, 1, 1), periodic=True, box=1.0)
assert torch.isfinite(vol).all()
sol = SmoothingLengthSolver()
for kern in ("lockstep", "diverging"):
    for box in (None, 1.0):
        hh = sol.solve(d(pos), 24, box, kernel=kern)
        sub = sol.solve(d(pos[np.argsort(pos[:, 0])]), 24, box, q_begin=100, q_count=700, kernel=kern)
        lists = sol.solve(d(pos), 9, box, want_neighbours=True, want_distances=True, kernel=kern)
        assert torch.isfinite(hh).all() and torch.isfinite(sub).all()
t = IonisationTableBase(-np.abs(rng.normal(size=(9, 11, 5))), np.linspace(-8, 2, 9), np.linspace(2, 9, 11), np.linspace(0, 9, 5), redshift_input_index=2)
v = t.evaluate_at_redshift(rng.uniform([-9, 1], [3, 10], (5000, 2)), 2.2)
torch.cuda.synchronize()
print("all_entry_points_small: ok")
