#!/usr/bin/env python3
"""One rank's share of a multi-GPU k-NN (queries = an index range of a lattice-ordered set): reach-limited against full build."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pos, _ = synthetic.s1_positions(n)
pos_d = torch.from_numpy(pos).cuda()
N = pos.shape[0]
sol = SmoothingLengthSolver()
for g in (0, G // 2, G - 1):
    lo, hi = g * N // G, (g + 1) * N // G
    res = {}
    for full in (True, False):
        h = sol.solve(pos_d, 48, 1.0, q_begin=lo, q_count=hi - lo, full_build=full)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            h = sol.solve(pos_d, 48, 1.0, q_begin=lo, q_count=hi - lo, full_build=full)
        e1.record(); torch.cuda.synchronize()
        res["full_build" if full else "reach_limited"] = e0.elapsed_time(e1) / 3
        res["equal"] = bool(torch.equal(h, res.setdefault("_h", h)))
    res.pop("_h")
    print(json.dumps({"n": n, "rank": g, "of": G, **res}), flush=True)
