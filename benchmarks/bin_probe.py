"""probe: sub-pixel regime (every particle deposited by the binning kernel) against AST_BIN_STAGES / AST_BIN_MINB.
One subprocess per setting.  S1 n^3 -> (8n)^2, h = d_48 / 64, one weight field."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from astro_sph_tools_b200 import synthetic, CoordinateAxes
from astro_sph_tools_b200.tools.projections import Projector2D
n = int(sys.argv[1]); npix = 8 * n
pos_d = torch.empty((n ** 3, 3), dtype=torch.float64, device="cuda")
for i0, i1, blk in synthetic.s1_blocks(n):
    pos_d[i0:i1].copy_(torch.from_numpy(blk))
N = n ** 3
h = torch.full((N,), synthetic.s1_h_lattice_estimate(n, 48) / 64, dtype=torch.float64, device="cuda")
m = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
eng = Projector2D()
out = torch.empty((1, npix, npix), dtype=torch.float64, device="cuda")
f = lambda: eng.project(pos_d, h, [m], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
eng.project(pos_d, h, [m], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out, timing=True)
alg = N * 40 + npix * npix * 8
print("stages", os.environ.get("AST_BIN_STAGES"), "minb", os.environ.get("AST_BIN_MINB"), "n", n, "ms", round(ms, 4), "hbm_frac", round(alg / (ms * 1e-3) / 1e9 / 6554.2, 4),
      "bin_ms", round(eng.last_stats["stage_ms"][0], 4), "sum", float(out.sum()) / npix ** 2)
''' % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "512"
for st, mb in (("2", "3"), ("2", "4"), ("2", "5"), ("2", "6"), ("4", "4")):
    env = dict(os.environ, AST_BIN_STAGES=st, AST_BIN_MINB=mb)
    subprocess.run([sys.executable, "-c", CHILD, n], env=env, check=False)
