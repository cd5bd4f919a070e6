"""one short run of the sub-pixel regime (every particle deposited by the binning kernel): the command profiled with ncu.
S1 n^3 -> (8n)^2, h = d_48 estimate / 64, one weight field; prints the time of 5 calls."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic, CoordinateAxes
from astro_sph_tools_b200.tools.projections import Projector2D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0 / 64
npix = 8 * n
pos_d = torch.empty((n ** 3, 3), dtype=torch.float64, device="cuda")
for i0, i1, blk in synthetic.s1_blocks(n):
    pos_d[i0:i1].copy_(torch.from_numpy(blk))
N = n ** 3
h = torch.full((N,), synthetic.s1_h_lattice_estimate(n, 48) * scale, dtype=torch.float64, device="cuda")
m = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
eng = Projector2D()
out = torch.empty((1, npix, npix), dtype=torch.float64, device="cuda")
f = lambda: eng.project(pos_d, h, [m], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize()
print("n", n, "h_scale", scale, "ms", round(e0.elapsed_time(e1) / 5, 4), "pairs", eng.last_stats["n_pairs"], "sum", float(out.sum()) / npix ** 2)
