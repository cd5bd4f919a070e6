#!/usr/bin/env python3
"""Large-h stress: config 2 (S1 256^3 -> 2048^2, h = d_48) with a fraction of the particles at 16 x h, so
that they cover ~1000 tiles each and go through the large-h split kernel.  Pixel updates per particle grow 256-fold for
those particles, so the fair comparison is the time per pixel update."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic, CoordinateAxes
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D

n, npix = 256, 2048
pos, _ = synthetic.s1_positions(n)
N = len(pos)
pos_d = torch.from_numpy(pos).cuda()
h0 = SmoothingLengthSolver().solve(pos_d, 48, 1.0)
m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
eng = Projector2D()
out = torch.empty((1, npix, npix), dtype=torch.float64, device="cuda")
base_ms = None
for frac in (0.0, 1e-4, 1e-3, 1e-2):
    h = h0.clone()
    if frac > 0:
        sel = torch.from_numpy(np.random.default_rng(1).choice(N, int(N * frac), replace=False)).cuda()
        h[sel] *= 16.0
    f = lambda: eng.project(pos_d, h, [m_d], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    updates = float((np.pi * (2.0 * h * npix) ** 2).sum().item())          # pixels inside the supports (map edges ignored)
    eng.project(pos_d, h, [m_d], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out, timing=True)
    st = eng.last_stats
    rec = {"fraction_at_16h": frac, "ms": round(ms, 2), "n_huge": st["n_huge"], "n_pairs": st["n_pairs"], "rounds": st["n_rounds"],
           "pixel_updates": updates, "ns_per_1000_updates": round(ms * 1e6 / (updates / 1e3), 3), "stage_ms": [round(x, 2) for x in st["stage_ms"]]}
    if base_ms is None:
        base_ms = rec["ns_per_1000_updates"]
    rec["cost_per_update_vs_no_huge"] = round(rec["ns_per_1000_updates"] / base_ms, 3)
    print(json.dumps(rec), flush=True)
