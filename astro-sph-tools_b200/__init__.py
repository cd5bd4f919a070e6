"""B200-native SPH deposition (2-D projection maps, 3-D voxel grids) and smoothing-length k-NN.

Drop-in for the hot path of QuasarX1/astro-sph-tools: ``tools.projections.create_image`` and
``quartic_spline_kernel`` keep the reference's signatures (tools/projections/_projector.py:75-87,
_kernels.pyx:9).  All compute runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/astro_sph_b200.h); there is no CPU fallback.
"""
from ._CoordinateAxes import CoordinateAxes  # noqa: F401

__version__ = "0.1.0"
