"""one short run of the k-NN (the command profiled with ncu): S1 n^3, k = 48, periodic; prints the time of 3 solves."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos, _ = synthetic.s1_positions(n)
pos_d = torch.from_numpy(pos).cuda()
sol = SmoothingLengthSolver()
h = sol.solve(pos_d, 48, 1.0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): h = sol.solve(pos_d, 48, 1.0)
e1.record(); torch.cuda.synchronize()
print("n", n, "ms", round(e0.elapsed_time(e1) / 3, 3), "checksum", float(h.sum()))
