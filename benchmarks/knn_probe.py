"""probe: k-NN time vs cell_target (mean particles per cell)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
pos, _ = synthetic.s1_positions(256)
pos_d = torch.from_numpy(pos).cuda()
ref = None
for ct in (0.5, 1.0, 2.0, 4.0, 8.0, 16.0):
    sol = SmoothingLengthSolver(cell_target=ct)
    h = sol.solve(pos_d, 48, 1.0); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(2): h = sol.solve(pos_d, 48, 1.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 2
    if ref is None: ref = h.clone()
    print(f"cell_target {ct:5.1f}: {dt*1e3:7.1f} ms  equal={bool(torch.equal(h, ref))}")
