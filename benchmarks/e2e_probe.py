"""probe: end-to-end time of project_host vs batch size, and the raw pinned H2D rate"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astro_sph_tools_b200 import synthetic, CoordinateAxes
from astro_sph_tools_b200.tools.projections import Projector2D
from astro_sph_tools_b200.tools.smoothing import compute_smoothing_lengths_device
n = 256
pos, rng = synthetic.s1_positions(n); N = len(pos)
m = np.full(N, 1.0 / N); mT = m * 10 ** rng.uniform(4, 7, N)
h = compute_smoothing_lengths_device(torch.from_numpy(pos).cuda(), 48, 1.0).cpu().numpy()
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
pos, m, mT, h = pin(pos), pin(m), pin(mT), pin(h)
print("is_pinned seen by torch:", torch.from_numpy(pos[100:200]).is_pinned())
d = torch.empty((N, 3), dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3): d.copy_(torch.from_numpy(pos), non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
print(f"H2D pos 403 MB: {dt*1e3:.2f} ms = {403/dt/1e3:.1f} GB/s")
eng = Projector2D()
for bp in (1 << 30, 1 << 23, 1 << 22, 1 << 21, 1 << 20):
    for _ in range(2): eng.project_host(pos, h, [m, mT], (2048, 2048), CoordinateAxes.Z, (0, 1, 0, 1), batch_particles=bp, return_device=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): eng.project_host(pos, h, [m, mT], (2048, 2048), CoordinateAxes.Z, (0, 1, 0, 1), batch_particles=bp, return_device=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(f"batch_particles {bp}: {eng.last_stats.get('n_batches', 1)} batches, {dt*1e3:.1f} ms (device result)")
t = time.perf_counter(); out = eng.project_host(pos, h, [m, mT], (2048, 2048), CoordinateAxes.Z, (0, 1, 0, 1)); print(f"with D2H to numpy: {(time.perf_counter()-t)*1e3:.1f} ms")
