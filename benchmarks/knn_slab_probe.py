"""probe: k-NN on ONE slab of an 8-slab decomposition (x in [0, 1/8) + ghost zones of the 512^3 S1 set), as every rank of an
8-GPU job sees it, against the cell size: default (2 particles per cell if the set filled the box) vs 2 x fill fraction."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = 1.5 * synthetic.s1_h_lattice_estimate(n, 48)
own, gh = [], []
for i0, i1, blk in synthetic.s1_blocks(n):
    x = blk[:, 0]
    o = x < 0.125
    g = (~o) & ((x < 0.125 + w) | (x > 1.0 - w))
    if o.any(): own.append(blk[o])
    if g.any(): gh.append(blk[g])
own = np.concatenate(own); gh = np.concatenate(gh)
pos = torch.from_numpy(np.ascontiguousarray(np.concatenate([own, gh]))).cuda()
nq = len(own)
fill = 0.125 + 2 * w
sol = SmoothingLengthSolver()
ref = None
for ct in (None, 2.0 * fill, 1.0 * fill, 4.0 * fill):
    f = lambda: sol.solve(pos, 48, 1.0, q_begin=0, q_count=nq, cell_target=ct)
    h = f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2): h = f()
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = h.clone()
    print(f"local set {len(pos)} (owned {nq}), fill {fill:.3f}, cell_target {ct}: {e0.elapsed_time(e1) / 2:.1f} ms, equal {bool(torch.equal(h, ref))}")
