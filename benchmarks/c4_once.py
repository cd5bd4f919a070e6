"""one short run of config 4's 3-D gridding (the command profiled with ncu): NFW-clustered 256^3 -> 512^3 voxels, periodic."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Gridder3D
import bench
N, ng = 256 ** 3, 512
pos = bench.nfw_positions_device(torch, torch.device("cuda"), N)
sol = SmoothingLengthSolver()
h = sol.solve(pos, 48, 1.0)
sol._ws = None
m = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
g = Gridder3D()
out = torch.empty((ng,) * 3, dtype=torch.float64, device="cuda")
for _ in range(2):
    g.grid(pos, h, m, (ng,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out, timing=True)
print("stage_ms", [round(x, 2) for x in g.last_stats["stage_ms"]], "mass", float(out.sum()) / ng ** 3)
