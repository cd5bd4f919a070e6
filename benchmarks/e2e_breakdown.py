#!/usr/bin/env python3
"""Where the end-to-end time of config 2 goes: device-resident pass, batched host pipeline without the read-back, with it."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D

n = 256
pos, rng = synthetic.s1_positions(n)
N = pos.shape[0]
pos_d = torch.from_numpy(pos).cuda()
h_d = SmoothingLengthSolver().solve(pos_d, 48, 1.0)
m = np.full(N, 1.0 / N); mT = m * 10 ** rng.uniform(4, 7, N)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
posp, hp, mp, mTp = pin(pos), pin(h_d.cpu().numpy()), pin(m), pin(mT)
m_d, mT_d = torch.from_numpy(m).cuda(), torch.from_numpy(mT).cuda()
eng = Projector2D()
size = (8 * n, 8 * n); b = (0.0, 1.0, 0.0, 1.0)


def timeit(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


out = torch.empty((2,) + size, dtype=torch.float64, device="cuda")
res = {"device_resident": timeit(lambda: eng.project(pos_d, h_d, [m_d, mT_d], size, 2, b, out=out))}
for nb in (1 << 24, 1 << 23, 1 << 22):
    for ramp in (False, True):
        res[f"host_no_readback_b{nb >> 20}M_ramp{int(ramp)}"] = timeit(lambda: eng.project_host(posp, hp, [mp, mTp], size, 2, b, batch_particles=nb, ramp=ramp, return_device=True))
res["host_with_readback_default"] = timeit(lambda: eng.project_host(posp, hp, [mp, mTp], size, 2, b))
h2d = torch.empty(N * 6, dtype=torch.float64, device="cuda")
src = torch.from_numpy(np.concatenate([posp.ravel(), hp, mp, mTp])).pin_memory()
res["h2d_805MB_alone"] = timeit(lambda: h2d.copy_(src, non_blocking=True))
host = torch.empty((2,) + size, dtype=torch.float64, pin_memory=True)
res["d2h_67MB_alone"] = timeit(lambda: host.copy_(out))
print(json.dumps(res, indent=1))

# batching cost without any copy: the same batch boundaries on device-resident slices
def batched_device(cuts):
    for i in range(len(cuts) - 1):
        lo, hi = cuts[i], cuts[i + 1]
        eng.project(pos_d[lo:hi], h_d[lo:hi], [m_d[lo:hi], mT_d[lo:hi]], size, 2, b, out=out, accumulate=i > 0)
bn = N // 4
res2 = {"device_4_equal_batches": timeit(lambda: batched_device([0, bn, 2 * bn, 3 * bn, N])),
        "device_ramped_6_batches": timeit(lambda: batched_device([0, bn // 4, bn // 4 + bn // 2, bn // 4 + bn // 2 + bn, bn // 4 + bn // 2 + 2 * bn, bn // 4 + bn // 2 + 3 * bn, N])),
        "device_2_equal_batches": timeit(lambda: batched_device([0, N // 2, N]))}
print(json.dumps(res2, indent=1))
