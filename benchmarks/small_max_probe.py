#!/usr/bin/env python3
"""Direct-deposit threshold (small_max_px: largest pixel bbox area deposited by the binning kernel) against the footprint."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D

n = 256
pos, rng = synthetic.s1_positions(n)
N = len(pos)
pos_d = torch.from_numpy(pos).cuda()
h_d = SmoothingLengthSolver().solve(pos_d, 48, 1.0)
m = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda"); mT = m * 3.0
out = torch.empty((2, 2048, 2048), dtype=torch.float64, device="cuda")
for sc in (1 / 32, 1 / 16, 1 / 8, 1 / 4):
    row = {"h_scale": sc, "radius_px": float(2 * (h_d * sc).mean().item() * 2048)}
    ref = None
    for smx in (16, 36, 64, 144):
        eng = Projector2D(small_max_px=smx)
        hs = h_d * sc
        for _ in range(2):
            eng.project(pos_d, hs, [m, mT], (2048, 2048), 2, (0.0, 1.0, 0.0, 1.0), out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.project(pos_d, hs, [m, mT], (2048, 2048), 2, (0.0, 1.0, 0.0, 1.0), out=out)
        e1.record(); torch.cuda.synchronize()
        row[f"ms_smx{smx}"] = round(e0.elapsed_time(e1) / 3, 3)
        if ref is None:
            ref = out.clone()
        else:
            row[f"rel_smx{smx}"] = float(((out - ref).norm() / ref.norm()).item())
    print(json.dumps(row), flush=True)
