#!/usr/bin/env python3
"""k-NN (k = 48, periodic, S1 n^3) time against mean particles per cell, for the lock-step and the diverging query kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos, _ = synthetic.s1_positions(n)
pos_d = torch.from_numpy(pos).cuda()
ref = None
for kern in ("lockstep", "diverging"):
    for ct in (1.0, 2.0, 3.0, 4.0, 8.0, 16.0):
        sol = SmoothingLengthSolver(cell_target=ct)
        h = sol.solve(pos_d, 48, 1.0, kernel=kern)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            h = sol.solve(pos_d, 48, 1.0, kernel=kern)
        e1.record(); torch.cuda.synchronize()
        ref = h if ref is None else ref
        print(json.dumps({"kernel": kern, "cell_target": ct, "ms": e0.elapsed_time(e1) / 2, "equal": bool(torch.equal(h, ref))}), flush=True)
