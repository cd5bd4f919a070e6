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
    part = eng.project_host(positions, smoothing_lengths, props, image_size, projection_axis, (x_min, x_max, y_min, y_max),
                            kernel, periodic, box_size, return_device=True)
    part = reduce_maps(part, dst, group, all_ranks)
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if all_ranks or rank == dst:
        host = torch.empty(part.shape, dtype=part.dtype, pin_memory=True)
        host.copy_(part)
        return host.numpy()
    return None
