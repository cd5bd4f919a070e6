#!/usr/bin/env python3
"""End-to-end probe of config 2 through project_host with different batch schedules (pinned host arrays in, numpy maps out)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
from astro_sph_tools_b200.tools.projections import Projector2D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos, rng = synthetic.s1_positions(n)
N = pos.shape[0]
h = SmoothingLengthSolver().solve(torch.from_numpy(pos).cuda(), 48, 1.0).cpu().numpy()
m = np.full(N, 1.0 / N); mT = m * 10 ** rng.uniform(4, 7, N)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
pos, h, m, mT = pin(pos), pin(h), pin(m), pin(mT)
eng = Projector2D()
size = (8 * n, 8 * n)
for sched in sys.argv[2:] or ["22", "23", "24"]:
    kw = {"batch_particles": 1 << int(sched)} if sched.isdigit() else {"schedule": sched}
    f = lambda: eng.project_host(pos, h, [m, mT], size, 2, (0.0, 1.0, 0.0, 1.0), **kw)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"schedule": sched, "ms_median": float(np.median(ts)), "ms_min": min(ts), "batches": eng.last_stats.get("n_batches", 1)}))
