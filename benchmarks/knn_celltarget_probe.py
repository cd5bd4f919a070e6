"""probe: mean particles per cell of the k-NN grid (S1 256^3, k = 48, periodic) -- the selection kernel sweeps fewer candidates on
smaller cells but more queries reach beyond its two cells of reach and fall back to the lock-step kernel."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
pos, _ = synthetic.s1_positions(256)
pos_d = torch.from_numpy(pos).cuda()
ref = None
for ct in (2.0, 1.5, 1.6, 1.75, 1.9, 2.2, 2.5):
    sol = SmoothingLengthSolver(cell_target=ct)
    h = sol.solve(pos_d, 48, 1.0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): h = sol.solve(pos_d, 48, 1.0)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = h.clone()
    print(json.dumps({"cell_target": ct, "ms": round(e0.elapsed_time(e1) / 3, 3), "equal": bool(torch.equal(h, ref))}), flush=True)
    del sol
