"""Ionisation-table lookup on the GPU and the fused ion column-density weight.

Mirror of the reference's ``IonisationTableBase`` (data_structures/_IonisationTable.py:30-69): same constructor
``(table, *table_positions, redshift_input_index=-1)``, ``__call__(gas_state)``, ``evaluate_at_redshift(gas_state,
redshift)`` and accessors.  The reference delegates to ``scipy.interpolate.RegularGridInterpolator(table_positions, table,
bounds_error=False, fill_value=-inf)`` (:44-49); here the same arithmetic runs in a CUDA kernel (ast_table_interp) and is
bit-equal to scipy in float64.  A concrete table (``IonisationTable_HM01``, io/ionisation_tables/_HM01.py:60-97) is this
class constructed from the arrays of its HDF5 file: ``(ionbal, logd, logt, redshift, redshift_input_index=2)``; reading the
file is I/O and out of scope.

``ion_weights`` / ``create_ion_image`` fuse the step that PRODUCES ``particle_properties`` for an ion column-density map
(element mass x 10**log10-ion-fraction) with the deposition: the weights are evaluated on the device and handed to the
projection kernels without an N-sized round trip through host memory.

CUDA only; there is no CPU fallback.

One documented deviation: ``redshift_input_index < 0`` counts from the end (``ndim + index``).  The reference computes
``ndim - index`` (:33), which is past the last dimension and makes ``evaluate_at_redshift`` raise IndexError for the
default -1; every concrete table passes a non-negative index.
"""
import ctypes as C

import numpy as np

from ... import _lib


class IonisationTableBase:
    def __init__(self, table, *table_positions, redshift_input_index: int = -1, device=None):
        n = len(table_positions)
        if n == 0:
            raise IndexError("No input dimensions were specified for table interpolation construction.")
        table = np.asarray(table)
        if len(table.shape) != n:
            raise IndexError(f"Interpolation table has {len(table.shape)} dimensions but {n} arrays were used to specify the table positions.")
        if n > _lib.TABLE_MAX_DIM:
            raise NotImplementedError(f"at most {_lib.TABLE_MAX_DIM} table dimensions are supported on the device, got {n}")
        self.__n = n
        self.__zdim = redshift_input_index if redshift_input_index >= 0 else n + redshift_input_index
        self.__axes = tuple(np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in table_positions)
        self.__table = np.ascontiguousarray(table, dtype=np.float64)
        for d, a in enumerate(self.__axes):
            if a.ndim != 1 or a.shape[0] != self.__table.shape[d]:
                raise ValueError(f"There are {a.shape[0] if a.ndim == 1 else a.shape} points and {self.__table.shape[d]} values in dimension {d}")
            if a.shape[0] < 2:
                raise NotImplementedError("dimensions of length one are not supported")
            if not np.all(a[1:] > a[:-1]):
                raise ValueError(f"The points in dimension {d} must be strictly ascending")     # scipy raises the same for unsorted grids
        self.__device = device
        self.__dev = None

    # ---- device state (uploaded on first use) ----------------------------------------------------------
    def _device_state(self):
        if self.__dev is None:
            torch = _lib.require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device()) if self.__device is None else torch.device(self.__device)
            self.__dev = dict(torch=torch, device=dev, lib=_lib.load(), table=torch.from_numpy(self.__table).to(dev),
                              axes=[torch.from_numpy(a).to(dev) for a in self.__axes])
        return self.__dev

    def _params(self, fixed_dim, fixed_value, pow10, fill_value=-np.inf):
        st = self._device_state()
        p = _lib.TableParams()
        p.ndim = self.__n
        for d in range(self.__n):
            p.shape[d] = self.__table.shape[d]
            p.axes[d] = st["axes"][d].data_ptr()
        p.table = st["table"].data_ptr()
        p.fill_value = float(fill_value)
        p.fixed_dim = int(fixed_dim)
        p.fixed_value = float(fixed_value)
        p.flags = _lib.TABLE_POW10 if pow10 else 0
        return p

    def device_eval(self, columns, redshift=None, base=None, pow10=False, out=None, stream=None):
        """Device-resident evaluation.  ``columns``: one float64 CUDA tensor per table dimension (any stride; views into an
        (N,d) tensor are fine), with ``None`` at the redshift dimension when ``redshift`` is given.  Returns a float64 CUDA
        tensor (N,): the interpolated value, ``10**value`` with pow10, times ``base`` when given."""
        st = self._device_state()
        torch = st["torch"]
        fixed = self.__zdim if redshift is not None else -1
        if len(columns) != self.__n:
            raise ValueError(f"The requested sample points xi have dimension {len(columns)} but this table has dimension {self.__n}")
        n = None
        ptrs = (C.c_void_p * _lib.TABLE_MAX_DIM)()
        strides = (C.c_int64 * _lib.TABLE_MAX_DIM)()
        for d, c in enumerate(columns):
            if d == fixed:
                continue
            if c is None or c.dtype != torch.float64 or c.ndim != 1 or not c.is_cuda:
                raise ValueError("coordinate columns must be 1-D float64 CUDA tensors")
            n = c.shape[0] if n is None else n
            if c.shape[0] != n:
                raise ValueError("coordinate columns must have the same length")
            ptrs[d] = c.data_ptr()
            strides[d] = c.stride(0)
        if n is None:
            raise ValueError("no coordinate column given")
        if base is not None and (base.dtype != torch.float64 or base.shape != (n,) or not base.is_cuda or not base.is_contiguous()):
            raise ValueError("base must be a contiguous float64 CUDA tensor (N,)")
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=st["device"])
        p = self._params(fixed, redshift if redshift is not None else 0.0, pow10)
        with torch.cuda.device(st["device"]):
            _lib.check(st["lib"].ast_table_interp(C.byref(p), ptrs, strides, C.c_int64(n), _lib.ptr(base), _lib.ptr(out),
                                                  _lib.stream_ptr(stream)))
        return out

    # ---- the reference's interface (numpy in, numpy out) -------------------------------------------------
    def _host_eval(self, gas_state, redshift):
        st = self._device_state()
        torch = st["torch"]
        gs = np.asarray(gas_state, dtype=float)
        want = self.__n - (1 if redshift is not None else 0)
        if gs.ndim == 1:
            gs = gs.reshape(1, -1) if gs.shape[0] == want else gs.reshape(-1, 1)
        lead = gs.shape[:-1]
        if gs.shape[-1] != want:
            raise ValueError(f"The requested sample points xi have dimension {gs.shape[-1]} but this RegularGridInterpolator has dimension {want}")
        gs = np.ascontiguousarray(gs.reshape(-1, want))
        g = torch.from_numpy(gs).to(st["device"])
        cols, j = [], 0
        for d in range(self.__n):
            if redshift is not None and d == self.__zdim:
                cols.append(None)
            else:
                cols.append(g[:, j]); j += 1
        if g.shape[0] == 0:
            return np.empty(lead, dtype=float)
        return self.device_eval(cols, redshift=redshift).cpu().numpy().reshape(lead)

    def __call__(self, gas_state: np.ndarray) -> np.ndarray:
        return self._host_eval(gas_state, None)

    def evaluate_at_redshift(self, gas_state: np.ndarray, redshift: float) -> np.ndarray:
        if not 0 <= self.__zdim < self.__n:
            raise IndexError(f"index {self.__zdim} is out of bounds for axis 1 with size {self.__n}")
        return self._host_eval(gas_state, float(redshift))

    @property
    def number_of_input_dimensions(self) -> int:
        return self.__n

    @property
    def ionisation_fraction_table(self) -> np.ndarray:
        return self.__table.copy()

    def get_table_dimension(self, dimension: int) -> np.ndarray:
        return self.__axes[dimension].copy()


def ion_weights(table: IonisationTableBase, element_masses, log10_hydrogen_number_density, log10_temperature, redshift,
                return_device=False):
    """element_mass_i * 10**table(log10 nH_i, log10 T_i, z): the ``particle_properties`` of an ion column-density map.
    Arrays are (N,) float64 numpy (or CUDA tensors).  Particles outside the table get weight 0 (10**-inf)."""
    st = table._device_state()
    torch = st["torch"]
    up = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(st["device"], non_blocking=True)
    w = table.device_eval([up(log10_hydrogen_number_density), up(log10_temperature), None], redshift=float(redshift),
                          base=up(element_masses).contiguous(), pow10=True)
    return w if return_device else w.cpu().numpy()


def create_ion_image(table: IonisationTableBase, positions, smoothing_lengths, element_masses, log10_hydrogen_number_density,
                     log10_temperature, redshift, image_size, chunk_size, projection_axis, x_min, x_max, y_min, y_max,
                     kernel_func=None, *, periodic=False, box_size=None):
    """Ion column-density map in one device-resident pass: table lookup -> weights -> create_image's deposition.  Same
    trailing arguments as ``create_image`` (tools/projections/_projector.py:75-87); returns (nx, ny) float64 numpy."""
    from ..projections._kernels import kernel_id_of, quartic_spline_kernel
    from ..projections._projector import _validate, default_projector
    kernel = kernel_id_of(kernel_func if kernel_func is not None else quartic_spline_kernel)
    positions, smoothing_lengths, (m, lognh, logt) = _validate(positions, smoothing_lengths,
                                                               [element_masses, log10_hydrogen_number_density, log10_temperature])
    st = table._device_state()
    torch = st["torch"]
    eng = default_projector(st["device"])
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(st["device"], non_blocking=True)
    w = ion_weights(table, up(m), up(lognh), up(logt), redshift, return_device=True)
    out = eng.project(up(positions), up(smoothing_lengths), w, image_size, projection_axis, (x_min, x_max, y_min, y_max), kernel,
                      periodic, box_size)
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out)
    return host.numpy()
