"""Drop-in for astro_sph_tools.tools.projections (reference tools/projections/__init__.py:5-6)."""
from ._projector import create_image, create_images, default_projector
from ._kernels import (quartic_spline_kernel, wendland_c2_kernel, wendland_c2_kernel_3d, cubic_spline_kernel_2d,
                       kernel_id_of, TabulatedKernel)
from ._engine import Projector2D
from ._gridder import create_grid, Gridder3D, default_gridder
from ._ingest import strip_units, pinned, snapshot_maps
