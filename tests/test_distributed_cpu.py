"""Host-side multi-GPU logic on the CPU: world_size-2 gloo.  Particles shard by index, every rank produces a
partial map (here: the CPU oracle stands in for the per-rank CUDA deposit, which needs a GPU), the partial maps are
summed with the same `reduce_maps` collective bench.py and create_images_sharded use (NCCL on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_cloud, rel_l2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, all_ranks, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from astro_sph_tools_b200 import distributed as astd
    pos, h, prop = random_cloud(21, 3001, h_hi=0.8, signed=True)
    lo, hi = astd.shard_bounds(len(h), world, rank)
    part = oracle.project2d(pos[lo:hi], h[lo:hi], prop[lo:hi], (64, 64), 2, 0.0, 10.0, 0.0, 10.0)
    t = astd.reduce_maps(torch.from_numpy(part.copy()), dst=0, all_ranks=all_ranks)
    if rank == 0 or all_ranks:
        full = oracle.project2d(pos, h, prop, (64, 64), 2, 0.0, 10.0, 0.0, 10.0)
        q.put((rank, rel_l2(t.numpy(), full), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("all_ranks", [False, True])
def test_sharded_maps_sum_to_full_map(all_ranks):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, all_ranks, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = []
    while not q.empty():
        got.append(q.get())
    assert len(got) == (2 if all_ranks else 1)
    for rank, err, lo, hi in got:
        assert err < 1e-13


def test_shard_bounds_cover_everything_once():
    from astro_sph_tools_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_reduce_is_identity_without_process_group():
    from astro_sph_tools_b200.distributed import reduce_maps
    t = torch.arange(6, dtype=torch.float64)
    assert reduce_maps(t) is t


class _ScipySolver:
    """stands in for SmoothingLengthSolver on the CPU (the oracle's scipy call): only the sharding logic is under test"""
    def solve(self, pos, k, box_size=None, q_begin=0, q_count=0, **kw):
        import oracle
        assert set(kw) <= {"cell_target"}
        h = oracle.knn_scipy(pos.numpy(), k, box_size)[0]
        return torch.from_numpy(h[q_begin:q_begin + q_count] if q_count else h)


def _knn_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from astro_sph_tools_b200 import distributed as astd
    pos = np.random.default_rng(5).uniform(0, 1, (2003, 3))
    lo, hi = (0, 700) if rank == 0 else (700, 2003)                 # unequal shards: the all-gather pads
    h = astd.smoothing_lengths_sharded(torch.from_numpy(pos[lo:hi].copy()), 16, 1.0, solver=_ScipySolver())
    ref = oracle.knn_scipy(pos, 16, 1.0)[0]
    q.put((rank, bool(np.array_equal(h.numpy(), ref[lo:hi]))))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_smoothing_lengths_gather_and_slice():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(2))
    assert got == [(0, True), (1, True)]


def test_gather_positions_without_process_group():
    from astro_sph_tools_b200.distributed import gather_positions
    p = torch.zeros((5, 3), dtype=torch.float64)
    allp, off = gather_positions(p)
    assert allp is p and off == 0


# ---- sharded ID matching (the reference's ArrayReorder_MPI_2 contract, tools/_ArrayReorder.py:88-258) -----------------------
def _numpy_matcher(source_ids, target_ids, source_filter, target_filter):
    """stand-in for the GPU hash join: index of the equal source element per target, -1 where there is none"""
    order = np.argsort(source_ids, kind="stable")
    pos = np.searchsorted(source_ids[order], target_ids)
    pos[pos >= len(order)] = 0
    hit = (source_ids[order][pos] == target_ids) if len(order) else np.zeros(len(target_ids), dtype=bool)
    if target_filter is not None:
        hit &= np.asarray(target_filter, dtype=bool)
    return np.where(hit, order[pos] if len(order) else 0, -1).astype(np.int64)


def _numpy_row_gather(rows, index, default):
    out = np.empty((len(index),) + rows.shape[1:], dtype=rows.dtype)
    if default is not None:
        out[...] = default
    out[index >= 0] = rows[index[index >= 0]]
    return out


def _reorder_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from astro_sph_tools_b200 import distributed as astd
    from reorder_util import load_reorder
    g = load_reorder("filters")                                           # golden case produced by the reference's own class
    src, tgt, data = g["source_ids"], g["target_ids"], g["data"]
    sf, tf = g["source_order_filter"], g["target_order_filter"]
    s0, s1 = astd.shard_bounds(len(src), world, rank)
    cut = (len(tgt) * 3) // 10                                            # unequal target shards
    t0, t1 = (0, cut) if rank == 0 else (cut, len(tgt))
    r = astd.ShardedArrayReorder.create(src[s0:s1], tgt[t0:t1], sf[s0:s1], tf[t0:t1], matcher=_numpy_matcher, row_gather=_numpy_row_gather)
    out = r(data[s0:s1], default_value=g["default_value"])
    q.put((rank, bool(np.array_equal(out, g["result"][t0:t1])), bool(np.array_equal(r.target_filter, g["target_filter"][t0:t1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reorder_equals_the_reference_result_per_rank():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_reorder_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = []
    while not q.empty():
        got.append(q.get())
    assert sorted(g[0] for g in got) == [0, 1] and all(g[1] and g[2] for g in got)


# ---- slab + ghost-zone smoothing lengths ------------------------------------------------------------------
def _slab_worker(rank, world, port, case, q):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from astro_sph_tools_b200 import distributed as astd, synthetic
    if case == "lattice_periodic":
        pos, _ = synthetic.s1_positions(20); box = 1.0; k = 16
    elif case == "clustered_periodic":
        pos, _ = synthetic.s2_positions(6000, 1.0, n_haloes=4, seed=3); box = 1.0; k = 24
    else:                                                                  # open box, random order, unequal shards
        rng = np.random.default_rng(9); pos = rng.normal(size=(5000, 3)) * np.array([3.0, 1.0, 0.5]); box = None; k = 12
    n = len(pos)
    cut = (n * 2) // 5 if case == "open_unequal" else astd.shard_bounds(n, world, 0)[1]
    lo, hi = (0, cut) if rank == 0 else (cut, n)
    h, st = astd.smoothing_lengths_slabs(torch.from_numpy(pos[lo:hi].copy()), k, box, solver=_ScipySolver(), return_stats=True)
    ref = oracle.knn_scipy(pos, k, box)[0]
    q.put((rank, bool(np.array_equal(h.numpy(), ref[lo:hi])), st["iterations"], st["local_set"], n))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["lattice_periodic", "clustered_periodic", "open_unequal"])
def test_slab_ghost_smoothing_lengths_equal_the_full_search(case):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = []
    while not q.empty():
        got.append(q.get())
    assert sorted(g[0] for g in got) == [0, 1] and all(g[1] for g in got)
    if case == "lattice_periodic":                                        # positions are NOT replicated: slab + ghosts only
        assert all(g[3] < 0.95 * g[4] for g in got)
