"""probe: k-NN query kernel time against the cap on resident blocks per SM (AST_KNN_BLOCKS_PER_SM; 128 threads per block).
One subprocess per value (the library reads the variable once).  256^3 S1, k = 48, periodic."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
n = int(sys.argv[1])
pos, _ = synthetic.s1_positions(n)
pos_d = torch.from_numpy(pos).cuda()
sol = SmoothingLengthSolver()
h = sol.solve(pos_d, 48, 1.0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): h = sol.solve(pos_d, 48, 1.0)
e1.record(); torch.cuda.synchronize()
print("blocks_per_sm", os.environ.get("AST_KNN_BLOCKS_PER_SM"), "n", n, "ms", round(e0.elapsed_time(e1) / 3, 3), "checksum", float(h.sum()))
''' % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "256"
for v in ("0", "3", "4", "5", "6", "8", "10", "12"):
    env = dict(os.environ, AST_KNN_BLOCKS_PER_SM=v)
    subprocess.run([sys.executable, "-c", CHILD, n], env=env, check=False)
