"""probe: selection kernel (default) against the lock-step kernel alone -- S1 n^3 and the NFW-clustered set, k = 48, periodic.
AST_KNN_VERBOSE=1 prints how many queries the selection kernel left to the lock-step kernel."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sets = []
pos, _ = synthetic.s1_positions(n)
sets.append(("S1 %d^3" % n, torch.from_numpy(pos).cuda()))
sets.append(("NFW %d^3" % n, bench.nfw_positions_device(torch, torch.device("cuda"), n ** 3)))
sol = SmoothingLengthSolver()
for name, pos_d in sets:
    ref = None
    for kernel in ("lockstep", "select"):
        h = sol.solve(pos_d, 48, 1.0, kernel=kernel); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): h = sol.solve(pos_d, 48, 1.0, kernel=kernel)
        e1.record(); torch.cuda.synchronize()
        if ref is None: ref = h.clone()
        print(json.dumps({"set": name, "kernel": kernel, "ms": round(e0.elapsed_time(e1) / 3, 3), "equal_to_lockstep": bool(torch.equal(h, ref))}), flush=True)
