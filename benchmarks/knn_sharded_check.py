#!/usr/bin/env python3
"""torchrun check of distributed.smoothing_lengths_sharded: every rank holds an index range of S1 n^3, positions are
all-gathered over NCCL, each rank answers its range; rank 0 compares a sample with scipy and prints the time."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from astro_sph_tools_b200 import synthetic, distributed as astd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos, _ = synthetic.s1_positions(n)
lo, hi = astd.shard_bounds(len(pos), world, rank)
mine = torch.from_numpy(pos[lo:hi].copy()).cuda()
h = astd.smoothing_lengths_sharded(mine, 48, 1.0)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    h = astd.smoothing_lengths_sharded(mine, 48, 1.0)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
from scipy.spatial import cKDTree
sel = np.random.default_rng(rank).choice(hi - lo, 2000, replace=False)
ref = cKDTree(pos, boxsize=1.0).query(pos[lo:hi][sel], k=48, workers=-1)[0][:, 47]
ok = torch.tensor([int(np.array_equal(h.cpu().numpy()[sel], ref))], device="cuda"); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"config": f"smoothing_lengths_sharded S1 {n}^3, k=48 periodic, {world} GPUs (all-gather of positions + per-rank query)",
                      "ms": float(ms.item()), "bit_equal_to_scipy_on_2000_queries_per_rank": bool(ok.item())}))
dist.destroy_process_group()
