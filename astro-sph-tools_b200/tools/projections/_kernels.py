"""SPH kernel functions W(r, h) with the reference's call signature
(tools/projections/_kernels.pyx:9 ``quartic_spline_kernel(double[:] r, double[:] h)`` -> new float64 array).

The functions are usable on their own (they evaluate on the GPU through ast_kernel_eval, float64) and they are
the values ``create_image(kernel_func=...)`` recognises: each one maps to a device kernel id, because a
Python callable cannot be invoked from a CUDA kernel.  ``quartic_spline_kernel`` keeps the reference's
(misleading) name: it is the M4 cubic spline with 1/(pi h^3) normalisation and support r < 2h.
"""
import ctypes as C

import numpy as np

from ... import _lib


def _check_double_1d(a, name):
    """Reproduce the Cython typed-memoryview errors of the reference (double[:])."""
    a = np.asarray(a) if not isinstance(a, np.ndarray) else a
    if a.dtype != np.float64:
        got = {"float32": "float", "int64": "long", "int32": "int"}.get(a.dtype.name, a.dtype.name)
        raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{got}'")
    if a.ndim != 1:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 1, got {a.ndim})")
    return a


def _evaluate(kernel_name, r, h):
    r = _check_double_1d(r, "r")
    h = _check_double_1d(h, "h")
    if r.shape != h.shape:
        raise ValueError("r and h must have the same length")
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    r_d = torch.from_numpy(np.ascontiguousarray(r)).to(dev)
    h_d = torch.from_numpy(np.ascontiguousarray(h)).to(dev)
    out = torch.empty_like(r_d)
    _lib.check(lib.ast_kernel_eval(C.c_int(_lib.KERNEL_IDS[kernel_name]), _lib.ptr(r_d), _lib.ptr(h_d), _lib.ptr(out),
                                   C.c_int64(r.size), _lib.stream_ptr()))
    return out.cpu().numpy()


def quartic_spline_kernel(r, h):
    """Reference kernel (tools/projections/_kernels.pyx:9-20): q=r/h; q<1: (1-1.5q^2+0.75q^3)/(pi h^3);
    1<=q<2: 0.25(2-q)^3/(pi h^3); else 0."""
    return _evaluate("cubic_spline_3d", r, h)


def wendland_c2_kernel(r, h):
    """Wendland C2 with 2-D normalisation and H = 2h: 7/(pi H^2) (1-u)^4 (1+4u), u = r/H (surface density maps)."""
    return _evaluate("wendland_c2_2d", r, h)


def wendland_c2_kernel_3d(r, h):
    """Wendland C2 with 3-D normalisation and H = 2h: 21/(2 pi H^3) (1-u)^4 (1+4u)."""
    return _evaluate("wendland_c2_3d", r, h)


def cubic_spline_kernel_2d(r, h):
    """M4 cubic spline with 2-D normalisation 10/(7 pi h^2) (mass-conserving surface density)."""
    return _evaluate("cubic_spline_2d", r, h)


_KNOWN = {quartic_spline_kernel: "cubic_spline_3d", wendland_c2_kernel: "wendland_c2_2d",
          wendland_c2_kernel_3d: "wendland_c2_3d", cubic_spline_kernel_2d: "cubic_spline_2d"}


class TabulatedKernel:
    """A user ``kernel_func`` of the reference form ``W(r, h) = f(r/h) / h**dim`` (``_projector.py:86`` accepts any
    ``Callable[[r, h], w]``), sampled on q = r/h in [0, 2] -- the reference masks r < 2h (``_pixel_calculations.pyx:31``) --
    and evaluated on the device by linear interpolation (AST_KERNEL_TABLE).  The callable itself runs on the host, once."""

    N = 8192

    def __init__(self, kernel_func):
        q = np.linspace(0.0, 2.0, self.N + 1)
        one = np.ones_like(q)
        f = np.asarray(kernel_func(q.copy(), one.copy()), dtype=np.float64)
        if f.shape != q.shape or not np.all(np.isfinite(f)):
            raise NotImplementedError("kernel_func(r, h) must return a finite float array of the length of r")
        scale = np.abs(f).max()
        self.dim = None
        for dim in (3, 2):
            ok = True
            for hh in (0.37, 2.5):
                g = np.asarray(kernel_func(q * hh, one * hh), dtype=np.float64) * hh ** dim
                ok = ok and np.allclose(g, f, rtol=1e-9, atol=1e-12 * scale)
            if ok:
                self.dim = dim
                break
        if self.dim is None:
            raise NotImplementedError(
                "kernel_func is not of the self-similar form f(r/h) / h**2 or f(r/h) / h**3, so it cannot be tabulated for the "
                "device (a Python callable cannot run inside a CUDA kernel and there is no CPU fallback)")
        self.pairs = np.ascontiguousarray(np.stack([f[:-1], f[1:] - f[:-1]], axis=1).astype(np.float32))
        self._dev = {}

    def device_table(self, torch, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.pairs).to(device)
        return self._dev[key]


_TABULATED = {}


def kernel_id_of(kernel_func):
    """Device kernel for a kernel_func argument: the name of a built-in kernel (one of the callables above or its name),
    or a TabulatedKernel for any other callable of the form f(r/h)/h**dim."""
    if isinstance(kernel_func, TabulatedKernel):
        return kernel_func
    if isinstance(kernel_func, str):
        if kernel_func in _lib.KERNEL_IDS:
            return kernel_func
        raise NotImplementedError(f"unknown kernel name {kernel_func!r}; known: {sorted(_lib.KERNEL_IDS)}")
    name = _KNOWN.get(kernel_func)
    if name is not None:
        return name
    if getattr(kernel_func, "__name__", "") == "quartic_spline_kernel":        # the reference's own compiled function object
        return "cubic_spline_3d"
    if not callable(kernel_func):
        raise NotImplementedError("kernel_func must be callable or the name of a built-in kernel")
    if kernel_func not in _TABULATED:
        _TABULATED[kernel_func] = TabulatedKernel(kernel_func)
    return _TABULATED[kernel_func]
