"""Multi-GPU projection: particles shard by index across ranks (one process per GPU), every rank deposits its
shard onto a full-size partial map, and the partial maps are summed with ONE collective.

This mirrors how the reference distributes particles (each MPI rank reads a disjoint particle subset,
io/EAGLE/_SnapshotEAGLE.py:120-130); the reference has no map reduction because its projector is single-process.
Deposition is linear in the particles, so sum_g map(shard_g) == map(all particles) up to float summation order.
The collective is torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
import numpy as np


def shard_bounds(n, world_size, rank):
    """contiguous index range [lo, hi) of rank `rank` when n particles are split over world_size ranks"""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_maps(partial, dst=0, group=None, all_ranks=False):
    """Sum the per-rank partial maps (a torch tensor, in place).  all_ranks=False: result valid on `dst` only."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    if all_ranks:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return partial


def project_sharded(projector, pos, h, props, image_size, axis, bounds, kernel="cubic_spline_3d", periodic=False, box=None,
                    dst=0, group=None, all_ranks=False, out=None):
    """Device-resident sharded projection: this rank's particles in, summed map(s) out (valid on dst / all ranks)."""
    part = projector.project(pos, h, props, image_size, axis, bounds, kernel, periodic, box, out=out)
    return reduce_maps(part, dst, group, all_ranks)


def create_images_sharded(positions, smoothing_lengths, particle_properties, image_size, chunk_size, projection_axis,
                          x_min, x_max, y_min, y_max, kernel_func=None, *, periodic=False, box_size=None, dst=0, group=None,
                          all_ranks=False):
    """Host-buffer entry for one rank of a multi-GPU job: same arguments as create_images, but the arrays hold only
    THIS rank's particles.  Returns the summed (P,nx,ny) float64 numpy maps on `dst` (or every rank), else None."""
    import torch
    import torch.distributed as dist
    from .tools.projections._projector import _validate, default_projector
    from .tools.projections._kernels import kernel_id_of, quartic_spline_kernel
    kernel = kernel_id_of(kernel_func if kernel_func is not None else quartic_spline_kernel)
    positions, smoothing_lengths, props = _validate(positions, smoothing_lengths, list(particle_properties))
    eng = default_projector()
    from . import _lib
    parts = []
    for s0 in range(0, len(props), _lib.MAX_PROPS):              # more weight fields than one pass takes: several passes, like create_images
        parts.append(eng.project_host(positions, smoothing_lengths, props[s0:s0 + _lib.MAX_PROPS], image_size, projection_axis,
                                      (x_min, x_max, y_min, y_max), kernel, periodic, box_size, return_device=True))
    part = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
    part = reduce_maps(part, dst, group, all_ranks)
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if all_ranks or rank == dst:
        host = torch.empty(part.shape, dtype=part.dtype, pin_memory=True)
        host.copy_(part)
        return host.numpy()
    return None


def gather_positions(pos_local, group=None):
    """All ranks hold consecutive index ranges of the particle set (the reference's per-rank read,
    io/EAGLE/_SnapshotEAGLE.py:120-130): returns (all positions (N,3) on this rank's device, offset of this rank's range).
    ONE collective (all-gather; NCCL over NVLink on GPUs), shards may differ in length (padded to the longest)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return pos_local, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([pos_local.shape[0]], dtype=torch.int64, device=pos_local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(t.item()) for t in sizes]
    n_max = max(sizes)
    padded = pos_local
    if pos_local.shape[0] < n_max:
        padded = torch.zeros((n_max, 3), dtype=pos_local.dtype, device=pos_local.device)
        padded[:pos_local.shape[0]] = pos_local
    parts = [torch.empty((n_max, 3), dtype=pos_local.dtype, device=pos_local.device) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:m] for p, m in zip(parts, sizes)], dim=0), sum(sizes[:rank])


def smoothing_lengths_sharded(pos_local, k=32, box_size=None, group=None, solver=None):
    """Multi-GPU smoothing lengths: this rank's particles in (a (n_g,3) float64 tensor on its GPU), this rank's h out (n_g,).
    Positions are all-gathered once (1024^3 x 24 B = 25.8 GB fits every 180 GB GPU); every rank then answers its own index
    range on a cell list built from the particles within reach of it (ast_knn_h with q_begin / q_count).  `solver` is any
    object with SmoothingLengthSolver.solve's signature (the CPU tests of this host logic inject a scipy-based one)."""
    if solver is None:
        from .tools.smoothing import SmoothingLengthSolver
        solver = SmoothingLengthSolver()
    pos_all, offset = gather_positions(pos_local, group)
    n_local = pos_local.shape[0]
    if n_local == pos_all.shape[0]:
        return solver.solve(pos_all, k, box_size)
    return solver.solve(pos_all, k, box_size, q_begin=offset, q_count=n_local)


# ---- ID matching across ranks ---------------------------------------------------------------------------
def gather_rows(local, group=None):
    """All-gather of per-rank arrays that differ in their first dimension (padded to the longest; ONE data collective plus
    the sizes).  Returns (concatenation in rank order, list of per-rank row counts)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local, [local.shape[0]]
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(t.item()) for t in sizes]
    n_max = max(sizes)
    padded = local
    if local.shape[0] < n_max:
        padded = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:m] for p, m in zip(parts, sizes)], dim=0), sizes


class ShardedArrayReorder:
    """The reference's MPI reorder (tools/_ArrayReorder.py:88-258 ArrayReorder_MPI_2: every rank holds a piece of the source
    IDs / data and a piece of the target IDs; IDs and data are gathered on the root rank, matched there on one core, and the
    reordered rows scattered back) with the root bottleneck removed: the (filtered) source IDs are all-gathered once at
    create(), every rank joins its OWN target IDs against them on its GPU (hash join, ast_match_ids), and a call all-gathers the
    source rows and gathers this rank's rows locally (ast_gather_rows).  Same contract per rank: output[i] is the source row
    whose ID equals this rank's target ID i, `default_value` where no rank holds it.

    `matcher(source_ids, target_ids, source_filter, target_filter) -> int64 index of the source element per target (-1: none)`
    and `row_gather(rows, index, default) -> rows[index]` default to the CUDA entry points; the CPU tests of this host logic
    inject numpy stand-ins (there is no CPU fallback on the product path)."""

    def __init__(self, t2s, n_source_local, source_filter_local, group, row_gather):
        self._t2s = t2s                                   # index into the gathered FILTERED source rows, -1 = no match
        self._n_source_local = n_source_local
        self._source_filter_local = source_filter_local
        self._group = group
        self._row_gather = row_gather
        self.target_filter = t2s >= 0
        self.matched_items = int(self.target_filter.sum())

    @staticmethod
    def create(source_order, target_order, source_order_filter=None, target_order_filter=None, group=None, matcher=None,
               row_gather=None, device=None):
        import torch
        if source_order.dtype != target_order.dtype:          # tools/_ArrayReorder.py:229-231
            raise TypeError(f"Input source and target order arrays have different datatypes ({source_order.dtype} and "
                            f"{target_order.dtype}).\nThese input arrays MUST declare matching datatypes.")
        if matcher is None:
            from .tools._ArrayReorder import _match as matcher
        if row_gather is None:
            from .tools._ArrayReorder import gather_rows_device as row_gather
        sf = np.ones(source_order.shape[0], dtype=bool) if source_order_filter is None else np.asarray(source_order_filter, dtype=bool)
        ids_local = torch.from_numpy(np.ascontiguousarray(source_order[sf]).astype(np.int64, copy=False))
        if device is not None:
            ids_local = ids_local.to(device)
        ids_all, _ = gather_rows(ids_local, group)
        t2s = np.asarray(matcher(ids_all.cpu().numpy(), np.asarray(target_order), None, target_order_filter))
        return ShardedArrayReorder(t2s, source_order.shape[0], sf, group, row_gather)

    def __call__(self, source_data, default_value=None):
        import torch
        if source_data.shape[0] != self._n_source_local:
            raise IndexError("One or more ranks provided the wrong number of input data elements.")     # :147-148
        if default_value is None and self.matched_items != self._t2s.shape[0]:
            raise ValueError("More output elements expected than matches but no default value provided and no output target "
                             "array to write matches to.")
        rows = torch.from_numpy(np.ascontiguousarray(source_data[self._source_filter_local]))
        rows_all, _ = gather_rows(rows, self._group)
        return self._row_gather(rows_all.numpy(), self._t2s, default_value)
