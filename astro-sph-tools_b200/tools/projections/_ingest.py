"""Snapshot-array ingestion for the projection path: the step immediately before ``create_image``.

The reference's readers hand the hot path ``unyt_array`` objects -- positions ``[Mpc, (N,3), float64]``, smoothing lengths
``[Mpc, (N,)]``, masses ``[Msun, (N,)]``, temperatures ``[K, (N,)]`` (io/data_structures/_SnapshotBase.py:599-616, :618-637,
:708-725, :889-909) -- and the Cython memoryviews of the pixel routine read them through the buffer protocol, ignoring the
units.  This module does the same without importing unyt (``strip_units``), optionally page-locks the host arrays in place so
the batched host-to-device copies of ``Projector2D.project_host`` run at full PCIe rate and overlap the deposition
(``pinned``), and drives the snapshot accessors for the two maps of BASELINE.json config 2 (``snapshot_maps``).
"""
import contextlib

import numpy as np

from ._kernels import quartic_spline_kernel
from ._projector import create_images


def strip_units(a, dtype_ok=(np.float64,)):
    """ndarray view of a unyt_array-like (or anything array-like) without its units; no copy for ndarray subclasses."""
    if isinstance(a, np.ndarray):
        return a.view(np.ndarray)
    for attr in ("ndarray_view", "to_ndarray"):
        f = getattr(a, attr, None)
        if callable(f):
            return np.asarray(f())
    v = getattr(a, "value", None)
    return np.asarray(v if v is not None else a)


@contextlib.contextmanager
def pinned(*arrays):
    """Page-lock contiguous host arrays IN PLACE for the duration of the block (cudaHostRegister: no copy, unlike
    ``torch.Tensor.pin_memory``), so host-to-device copies from them are asynchronous DMA transfers."""
    from ... import _lib
    torch = _lib.require_cuda()
    rt = torch.cuda.cudart()
    done = []
    try:
        for a in arrays:
            a = strip_units(a)
            if a.size == 0:
                continue
            if not a.flags.c_contiguous:
                raise ValueError("only C-contiguous arrays can be page-locked in place")
            err = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            if int(err) != 0:
                raise RuntimeError(f"cudaHostRegister failed with error {int(err)}")
            done.append(a)
        yield
    finally:
        torch.cuda.synchronize()
        for a in done:
            rt.cudaHostUnregister(a.ctypes.data)


def snapshot_maps(snapshot, particle_type, image_size, projection_axis, x_min, x_max, y_min, y_max,
                  kernel_func=quartic_spline_kernel, *, temperature=True, use_proper_units=False, pin=True, periodic=False,
                  box_size=None, allow_float32=False):
    """Projected mass map (and mass-weighted temperature map) of one particle type of a reference snapshot object.

    ``snapshot`` is anything with the reference's accessor interface (``SnapshotBase``: ``get_positions(pt,
    use_proper_units)``, ``get_smoothing_lengths(pt, use_proper_units)``, ``get_masses(pt)``, ``get_temperatures(pt)``).
    Both weight fields (m, m*T) are deposited in ONE pass over the particles; the temperature map is their ratio where the
    mass map is non-zero.  Returns ``{"mass": (nx,ny) float64, "temperature": (nx,ny) float64}`` (host arrays)."""
    pos = strip_units(snapshot.get_positions(particle_type, use_proper_units))
    h = strip_units(snapshot.get_smoothing_lengths(particle_type, use_proper_units))
    m = strip_units(snapshot.get_masses(particle_type))
    props = [m]
    if temperature:
        props.append(m * strip_units(snapshot.get_temperatures(particle_type)))
    arrays = [np.ascontiguousarray(a) for a in (pos, h, *props)]
    ctx = pinned(*arrays) if pin else contextlib.nullcontext()
    with ctx:
        maps = create_images(arrays[0], arrays[1], arrays[2:], image_size, 0, projection_axis, x_min, x_max, y_min, y_max,
                             kernel_func, periodic=periodic, box_size=box_size, allow_float32=allow_float32)
    out = {"mass": maps[0]}
    if temperature:
        with np.errstate(invalid="ignore", divide="ignore"):
            out["temperature"] = np.where(maps[0] != 0, maps[1] / maps[0], 0.0)
    return out
