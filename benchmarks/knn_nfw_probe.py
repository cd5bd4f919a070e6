"""probe: k-NN (k = 48, periodic) on the NFW-clustered 256^3 set; AST_KNN_DENSE_N27 (27-cell count from which a pending query moves
to the fine grid; 0 = one level only) is read once per process, so one process per setting."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pos_d = bench.nfw_positions_device(torch, torch.device("cuda"), n ** 3)
sol = SmoothingLengthSolver()
h = sol.solve(pos_d, 48, 1.0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): h = sol.solve(pos_d, 48, 1.0)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"dense_n27": os.environ.get("AST_KNN_DENSE_N27", "default"), "ms": round(e0.elapsed_time(e1) / 3, 2), "checksum": float(h.sum())}), flush=True)
