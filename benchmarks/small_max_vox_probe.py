#!/usr/bin/env python3
"""3-D direct-deposit threshold (small_max_vox) on the NFW-clustered set of config 4 and on the uniform lattice."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Gridder3D

n = 256; N = n ** 3
sol = SmoothingLengthSolver()
sets = {"nfw": synthetic.s2_positions(N, 1.0, n_haloes=512, seed=12345)[0], "lattice": synthetic.s1_positions(n)[0]}
for name, pos in sets.items():
    pos_d = torch.from_numpy(pos).cuda()
    h_d = sol.solve(pos_d, 48, 1.0)
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    out = torch.empty((2 * n,) * 3, dtype=torch.float64, device="cuda")
    row = {"set": name}
    for smv in (27, 64, 125, 216):
        g = Gridder3D(small_max_vox=smv)
        for _ in range(2):
            g.grid(pos_d, h_d, m_d, (2 * n,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            g.grid(pos_d, h_d, m_d, (2 * n,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out)
        e1.record(); torch.cuda.synchronize()
        row[f"ms_smv{smv}"] = round(e0.elapsed_time(e1) / 2, 2)
    print(json.dumps(row), flush=True)
    del pos_d, h_d, m_d, out
    sol._ws = None; torch.cuda.empty_cache()
