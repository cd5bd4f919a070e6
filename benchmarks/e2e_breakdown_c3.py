#!/usr/bin/env python3
"""Where the end-to-end time of config 3 (134 M particles -> 4096^2, one GPU) goes: device-resident pass, batched host pipeline
with different batch sizes without and with the read-back, the same batch boundaries on device-resident slices (no copies)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D
from astro_sph_tools_b200.tools.projections._engine import batch_cuts

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = n ** 3
pos_d = torch.empty((N, 3), dtype=torch.float64, device="cuda")
posp = torch.empty((N, 3), dtype=torch.float64, pin_memory=True).numpy()
for i0, i1, blk in synthetic.s1_blocks(n):
    pos_d[i0:i1].copy_(torch.from_numpy(blk)); posp[i0:i1] = blk
sol = SmoothingLengthSolver()
h_d = sol.solve(pos_d, 48, 1.0)
sol._ws = None
hp = torch.empty(N, dtype=torch.float64, pin_memory=True); hp.copy_(h_d); hp = hp.numpy()
mp = torch.full((N,), 1.0 / N, dtype=torch.float64, pin_memory=True).numpy()
m_d = torch.from_numpy(mp).cuda()
eng = Projector2D()
size = (8 * n, 8 * n); b = (0.0, 1.0, 0.0, 1.0)


def timeit(f, reps=3):
    f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(ts)), 2)


out = torch.empty((1,) + size, dtype=torch.float64, device="cuda")
res = {"device_resident": timeit(lambda: eng.project(pos_d, h_d, [m_d], size, 2, b, out=out))}
print(json.dumps(res), flush=True)
for nb in (1 << 22, 1 << 23, 1 << 24, 1 << 25):
    for ramp in (True,):
        res[f"host_no_readback_batch{nb >> 20}M_ramp{int(ramp)}"] = timeit(lambda: eng.project_host(posp, hp, [mp], size, 2, b, batch_particles=nb, ramp=ramp, return_device=True))
        res[f"n_batches_{nb >> 20}M"] = eng.last_stats.get("n_batches")
res["host_with_readback_default"] = timeit(lambda: eng.project_host(posp, hp, [mp], size, 2, b))
host = torch.empty((1,) + size, dtype=torch.float64, pin_memory=True)
res["d2h_134MB_alone"] = timeit(lambda: host.copy_(out))
print(json.dumps(res, indent=1), flush=True)


def batched_device(cuts):
    for i in range(len(cuts) - 1):
        lo, hi = cuts[i], cuts[i + 1]
        eng.project(pos_d[lo:hi], h_d[lo:hi], [m_d[lo:hi]], size, 2, b, out=out, accumulate=i > 0)
res2 = {}
for nbat in (4, 8, 16):
    _, cuts = batch_cuts(N, nbat, True)
    res2[f"device_slices_{len(cuts) - 1}_ramped_batches"] = timeit(lambda: batched_device(cuts))
print(json.dumps(res2, indent=1))
