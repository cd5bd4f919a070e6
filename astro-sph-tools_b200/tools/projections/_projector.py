"""create_image: the reference's public projection function with its signature and return type
(tools/projections/_projector.py:75-87 -> np.ndarray (nx, ny) float64, img[xi, yi]).

Semantics (reference file:line):
  * pixel sizes (x_max-x_min)/nx, (y_max-y_min)/ny                       _projector.py:34-35
  * sample point = pixel LOWER corner x_min + xi*dx                      _pixel_calculations.pyx:13-14
  * in-plane columns X->(1,2), Y->(0,2), Z->(0,1)                        _pixel_calculations.pyx:20-28
  * contributor iff dx**2 + dy**2 < (2h)**2 ; value = sum prop*W(r,h)    _pixel_calculations.pyx:30-34
  * chunk_size only bounds the reference's temporaries; its results do not depend on it, so
    it is accepted and ignored here.
The reference derives both pixel sizes from image_size[0] inside the pixel routine (_pixel_calculations.pyx:11-12)
which breaks non-square images (all-zero maps); here nx != ny uses (y_max-y_min)/ny as _projector.py:35 intends.
Keyword-only extensions default to reference behaviour.
"""
from typing import Callable, Sequence

import numpy as np

from ..._CoordinateAxes import CoordinateAxes
from ._engine import Projector2D
from ._kernels import kernel_id_of, quartic_spline_kernel

_default = {}


def default_projector(device=None):
    """Process-wide Projector2D per device (keeps the workspace between create_image calls)."""
    from ... import _lib
    torch = _lib.require_cuda()
    key = torch.cuda.current_device() if device is None else torch.device(device).index
    if key not in _default:
        _default[key] = Projector2D(device=None if device is None else device)
    return _default[key]


def _check_buffer(a, ndim, name, allow_float32=False):
    """Same failures as the reference's Cython typed memoryviews (double[:, :] / double[:])."""
    if hasattr(a, "value") and not isinstance(a, np.ndarray):      # unyt_array-like without importing unyt
        a = a.value
    a = np.asarray(a) if not isinstance(a, np.ndarray) else a
    if a.ndim != ndim:
        raise ValueError(f"Buffer has wrong number of dimensions (expected {ndim}, got {a.ndim})")
    if a.dtype == np.float32 and allow_float32:
        return np.asarray(a)
    if a.dtype != np.float64:
        got = {"float32": "float", "int64": "long", "int32": "int"}.get(a.dtype.name, a.dtype.name)
        raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{got}'")
    return np.asarray(a)


def _validate(positions, smoothing_lengths, props, allow_float32=False):
    positions = _check_buffer(positions, 2, "positions", allow_float32)
    smoothing_lengths = _check_buffer(smoothing_lengths, 1, "smoothing_lengths", allow_float32)
    props = [_check_buffer(q, 1, "particle_properties", allow_float32) for q in props]
    n = positions.shape[0]
    if positions.shape[1] != 3:
        raise ValueError(f"positions must have shape (N, 3), got {positions.shape}")
    if smoothing_lengths.shape[0] != n or any(q.shape[0] != n for q in props):
        raise ValueError("positions, smoothing_lengths and particle_properties must have the same length")
    return positions, smoothing_lengths, props


def create_image(
    positions: np.ndarray,
    smoothing_lengths: np.ndarray,
    particle_properties: np.ndarray,
    image_size: tuple,
    chunk_size: int,
    projection_axis: CoordinateAxes,
    x_min: float,
    x_max: float,
    y_min: float,
    y_max: float,
    kernel_func: Callable = quartic_spline_kernel,
    *,
    periodic: bool = False,
    box_size=None,
    device=None,
    allow_float32: bool = False,
) -> np.ndarray:
    """SPH-kernel-weighted projection of particles onto a 2-D map (see module docstring).
    allow_float32=True (extension): float32 arrays, as stored on disk, are accepted, sent over PCIe as float32 and widened
    on the device -- the result equals passing ``a.astype(numpy.float64)``; the default rejects them like the reference."""
    kernel = kernel_id_of(kernel_func)
    positions, smoothing_lengths, props = _validate(positions, smoothing_lengths, [particle_properties], allow_float32)
    eng = default_projector(device)
    return eng.project_host(positions, smoothing_lengths, props[0], image_size, projection_axis,
                            (x_min, x_max, y_min, y_max), kernel, periodic, box_size)


def create_images(
    positions: np.ndarray,
    smoothing_lengths: np.ndarray,
    particle_properties: Sequence[np.ndarray],
    image_size: tuple,
    chunk_size: int,
    projection_axis: CoordinateAxes,
    x_min: float,
    x_max: float,
    y_min: float,
    y_max: float,
    kernel_func: Callable = quartic_spline_kernel,
    *,
    periodic: bool = False,
    box_size=None,
    device=None,
    allow_float32: bool = False,
) -> np.ndarray:
    """Extension: several weight arrays (e.g. mass and mass*T) deposited in ONE pass over the particles.
    Returns (P, nx, ny); row p equals create_image(..., particle_properties[p], ...)."""
    kernel = kernel_id_of(kernel_func)
    positions, smoothing_lengths, props = _validate(positions, smoothing_lengths, list(particle_properties), allow_float32)
    eng = default_projector(device)
    from ... import _lib
    outs = []
    for s in range(0, len(props), _lib.MAX_PROPS):
        outs.append(eng.project_host(positions, smoothing_lengths, props[s:s + _lib.MAX_PROPS], image_size, projection_axis,
                                     (x_min, x_max, y_min, y_max), kernel, periodic, box_size))
    return np.concatenate(outs, axis=0)
