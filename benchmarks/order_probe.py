"""probe: the same particle set in lattice order and in a fixed random order ("sort cost only shows on the latter"):
config 2 (S1 256^3 -> 2048^2, h = d_48, one field) stage times; and config 4's 3-D grid in its (random) recipe order against the
same set ordered by brick."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic, CoordinateAxes
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D, Gridder3D
import bench

def timed(f, reps=3):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n, npix = 256, 2048
pos, _ = synthetic.s1_positions(n)
N = len(pos)
pos_d = torch.from_numpy(pos).cuda()
sol = SmoothingLengthSolver()
h_d = sol.solve(pos_d, 48, 1.0)
m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
eng = Projector2D()
out = torch.empty((1, npix, npix), dtype=torch.float64, device="cuda")
perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
for name, p_, h_ in (("lattice order", pos_d, h_d), ("random order", pos_d[perm].contiguous(), h_d[perm].contiguous())):
    f = lambda: eng.project(p_, h_, [m_d], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out, presort="auto")
    ms = timed(f)
    eng.project(p_, h_, [m_d], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), out=out, timing=True, presort="auto")
    print(json.dumps({"config": "2-D 256^3 -> 2048^2, " + name, "ms": round(ms, 2), "stage_ms": [round(x, 2) for x in eng.last_stats["stage_ms"]]}), flush=True)
del out, eng
# 3-D: NFW set in recipe (random) order and ordered by brick key
pos3 = bench.nfw_positions_device(torch, torch.device("cuda"), N)
h3 = sol.solve(pos3, 48, 1.0)
sol._ws = None
g = Gridder3D()
ng = 512
out3 = torch.empty((ng,) * 3, dtype=torch.float64, device="cuda")
key = ((pos3[:, 0] * 64).long().clamp(0, 63) * 64 + (pos3[:, 1] * 64).long().clamp(0, 63)) * 64 + (pos3[:, 2] * 64).long().clamp(0, 63)
order = torch.argsort(key)
for name, p_, h_ in (("recipe (random) order", pos3, h3), ("ordered by 8^3-voxel brick", pos3[order].contiguous(), h3[order].contiguous())):
    f = lambda: g.grid(p_, h_, m_d, (ng,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out3)
    ms = timed(f, 2)
    g.grid(p_, h_, m_d, (ng,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0, out=out3, timing=True)
    print(json.dumps({"config": "3-D NFW 256^3 -> 512^3, " + name, "ms": round(ms, 2), "stage_ms": [round(x, 2) for x in g.last_stats["stage_ms"]]}), flush=True)
