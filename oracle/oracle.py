"""ctypes front-end of oracle/liboracle.so (the plain-C float64 restatement, sph_oracle.c) and of the
reference's own compiled path in oracle/_ref (build_ref.py).  TEST INFRASTRUCTURE ONLY.

Parity status: the reference has no tests/golden vectors for this path; the restatement is pinned
against the reference's own compiled code (tests/test_oracle_vs_reference.py, tests/golden/).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KERNEL_IDS = {"cubic_spline_3d": 0, "wendland_c2_2d": 1, "wendland_c2_3d": 2, "cubic_spline_2d": 3}

_dp = C.POINTER(C.c_double)


def build(force=False):
    """Compile liboracle.so if missing (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_kernel_eval.restype = C.c_double
        _LIB.orc_kernel_eval.argtypes = [C.c_int, C.c_double, C.c_double]
        _LIB.orc_bin2d.restype = C.c_int64
        _LIB.orc_bin3d.restype = C.c_int64
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def images(periodic=False, box=None):
    """Periodic image shifts in the fixed enumeration order m = 3*(ia+1) + (ib+1), ia,ib in {-1,0,1}.
    Returns (n_img, shift_a, shift_b).  box = (box_a, box_b) in-plane lengths."""
    if not periodic:
        return 1, np.zeros(1), np.zeros(1)
    ba, bb = box
    sa = np.array([ia * ba for ia in (-1, 0, 1) for ib in (-1, 0, 1)], dtype=np.float64)
    sb = np.array([ib * bb for ia in (-1, 0, 1) for ib in (-1, 0, 1)], dtype=np.float64)
    return 9, sa, sb


def images3(periodic=False, box=None):
    if not periodic:
        return 1, np.zeros(3)
    s = np.array([[ia * box[0], ib * box[1], ic * box[2]] for ia in (-1, 0, 1) for ib in (-1, 0, 1) for ic in (-1, 0, 1)],
                 dtype=np.float64)
    return 27, s


def kernel_eval(kernel, r, h):
    kid = KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
    r = _f64(np.atleast_1d(r)); h = _f64(np.atleast_1d(h))
    out = np.empty_like(r)
    lib().orc_kernel_eval_array(C.c_int(kid), _p(r), _p(h), _p(out), C.c_int64(r.size))
    return out


def _common(pos, h):
    pos = _f64(pos); h = _f64(h)
    assert pos.ndim == 2 and pos.shape[1] == 3 and h.shape == (pos.shape[0],)
    return pos, h, pos.shape[0]


def project2d(pos, h, prop, image_size, axis, x_min, x_max, y_min, y_max, kernel="cubic_spline_3d",
              periodic=False, box=None, form="scatter", nthreads=0):
    """Oracle map(s).  prop: (N,) -> (nx,ny) ; (P,N) -> (P,nx,ny).  form = 'scatter' | 'gather'."""
    pos, h, N = _common(pos, h)
    prop = _f64(prop)
    single = prop.ndim == 1
    prop2 = prop.reshape(1, N) if single else prop
    nx, ny = int(image_size[0]), int(image_size[1])
    kid = KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
    n_img, sa, sb = images(periodic, box)
    out = np.zeros((prop2.shape[0], nx, ny), dtype=np.float64)
    if form == "gather":
        for p in range(prop2.shape[0]):
            pr = np.ascontiguousarray(prop2[p])
            lib().orc_project2d_gather(_p(pos), _p(h), _p(pr), C.c_int64(N), C.c_int(axis), C.c_int(nx), C.c_int(ny),
                                       C.c_double(x_min), C.c_double(x_max), C.c_double(y_min), C.c_double(y_max),
                                       C.c_int(kid), C.c_int(n_img), _p(sa), _p(sb), _p(out[p]))
    else:
        lib().orc_project2d_scatter(_p(pos), _p(h), _p(prop2), C.c_int(prop2.shape[0]), C.c_int64(N), C.c_int(axis),
                                    C.c_int(nx), C.c_int(ny), C.c_double(x_min), C.c_double(x_max),
                                    C.c_double(y_min), C.c_double(y_max), C.c_int(kid), C.c_int(n_img), _p(sa), _p(sb),
                                    C.c_int(nthreads), _p(out))
    return out[0] if single else out


def bbox2d(pos, h, image_size, axis, x_min, x_max, y_min, y_max, periodic=False, box=None, brute=False):
    pos, h, N = _common(pos, h)
    nx, ny = int(image_size[0]), int(image_size[1])
    n_img, sa, sb = images(periodic, box)
    out = np.empty((n_img * N, 4), dtype=np.int32)
    lib().orc_bbox2d(_p(pos), _p(h), C.c_int64(N), C.c_int(axis), C.c_int(nx), C.c_int(ny), C.c_double(x_min),
                     C.c_double(x_max), C.c_double(y_min), C.c_double(y_max), C.c_int(n_img), _p(sa), _p(sb),
                     C.c_int(int(brute)), _p(out))
    return out


def contrib_count2d(pos, h, image_size, axis, x_min, x_max, y_min, y_max, periodic=False, box=None):
    pos, h, N = _common(pos, h)
    nx, ny = int(image_size[0]), int(image_size[1])
    n_img, sa, sb = images(periodic, box)
    out = np.empty((nx, ny), dtype=np.int32)
    lib().orc_contrib_count2d(_p(pos), _p(h), C.c_int64(N), C.c_int(axis), C.c_int(nx), C.c_int(ny), C.c_double(x_min),
                              C.c_double(x_max), C.c_double(y_min), C.c_double(y_max), C.c_int(n_img), _p(sa), _p(sb), _p(out))
    return out


def bin2d(pos, h, image_size, axis, x_min, x_max, y_min, y_max, tile=32, small_max_px=16, huge_min_tiles=256,
          periodic=False, box=None, brute=False, sort=True):
    """Index work: returns dict(cls, pairs (emit order), sorted (stable by key), huge)."""
    pos, h, N = _common(pos, h)
    nx, ny = int(image_size[0]), int(image_size[1])
    n_img, sa, sb = images(periodic, box)
    args = [_p(pos), _p(h), C.c_int64(N), C.c_int(axis), C.c_int(nx), C.c_int(ny), C.c_double(x_min), C.c_double(x_max),
            C.c_double(y_min), C.c_double(y_max), C.c_int(n_img), _p(sa), _p(sb), C.c_int(tile), C.c_int64(small_max_px),
            C.c_int64(huge_min_tiles), C.c_int(int(brute))]
    nh = C.c_int64(0)
    cls = np.zeros(n_img * N, dtype=np.uint8)
    npairs = lib().orc_bin2d(*args, _p(cls), None, None, C.byref(nh))
    pairs = np.empty(max(npairs, 1), dtype=np.uint64)
    huge = np.empty(max(nh.value, 1), dtype=np.uint64)
    lib().orc_bin2d(*args, _p(cls), _p(pairs), _p(huge), C.byref(nh))
    pairs = pairs[:npairs]; huge = huge[:nh.value]
    res = dict(cls=cls.reshape(n_img, N), pairs=pairs, huge=huge)
    if sort:
        s = pairs.copy(); tmp = np.empty_like(s)
        if npairs:      # stable by TILE key: the 4 image bits below it are not sorted on (pairs of a tile stay in emit order)
            lib().orc_sort_pairs_stable_bits(_p(s), C.c_int64(npairs), _p(tmp), C.c_int(4 if n_img > 1 else 0))
        res["sorted"] = s
    return res


def grid3d(pos, h, prop, grid_size, lo, hi, kernel="cubic_spline_3d", periodic=False, box=None, nthreads=0):
    pos, h, N = _common(pos, h)
    prop = _f64(prop)
    nx, ny, nz = (int(v) for v in grid_size)
    lo = _f64(lo); hi = _f64(hi)
    kid = KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
    n_img, s3 = images3(periodic, box)
    out = np.zeros((nx, ny, nz), dtype=np.float64)
    lib().orc_grid3d_scatter(_p(pos), _p(h), _p(prop), C.c_int64(N), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(lo), _p(hi),
                             C.c_int(kid), C.c_int(n_img), _p(np.ascontiguousarray(s3)), C.c_int(nthreads), _p(out))
    return out


def bbox3d(pos, h, grid_size, lo, hi, periodic=False, box=None):
    pos, h, N = _common(pos, h)
    nx, ny, nz = (int(v) for v in grid_size)
    lo = _f64(lo); hi = _f64(hi)
    n_img, s3 = images3(periodic, box)
    out = np.empty((n_img * N, 6), dtype=np.int32)
    lib().orc_bbox3d(_p(pos), _p(h), C.c_int64(N), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(lo), _p(hi), C.c_int(n_img),
                     _p(np.ascontiguousarray(s3)), _p(out))
    return out


def bin3d(pos, h, grid_size, lo, hi, brick=8, small_max_vox=64, huge_min_bricks=512, periodic=False, box=None, brute=False):
    """3-D index work: dict(cls, pairs (emit order), sorted (stable by key), huge)."""
    pos, h, N = _common(pos, h)
    nx, ny, nz = (int(v) for v in grid_size)
    lo = _f64(lo); hi = _f64(hi)
    n_img, s3 = images3(periodic, box)
    s3 = np.ascontiguousarray(s3)
    args = [_p(pos), _p(h), C.c_int64(N), C.c_int(nx), C.c_int(ny), C.c_int(nz), _p(lo), _p(hi), C.c_int(n_img), _p(s3),
            C.c_int(brick), C.c_int64(small_max_vox), C.c_int64(huge_min_bricks), C.c_int(int(brute))]
    nh = C.c_int64(0)
    cls = np.zeros(n_img * N, dtype=np.uint8)
    npairs = lib().orc_bin3d(*args, _p(cls), None, None, C.byref(nh))
    pairs = np.empty(max(npairs, 1), dtype=np.uint64)
    huge = np.empty(max(nh.value, 1), dtype=np.uint64)
    lib().orc_bin3d(*args, _p(cls), _p(pairs), _p(huge), C.byref(nh))
    pairs = pairs[:npairs]; huge = huge[:nh.value]
    s = pairs.copy(); tmp = np.empty_like(s)
    if npairs:          # stable by BRICK key: the 5 image bits below it are not sorted on
        lib().orc_sort_pairs_stable_bits(_p(s), C.c_int64(npairs), _p(tmp), C.c_int(5 if n_img > 1 else 0))
    return dict(cls=cls.reshape(n_img, N), pairs=pairs, sorted=s, huge=huge)


def knn_brute(pos, k, box=0.0, want_lists=False):
    pos = _f64(pos); N = pos.shape[0]
    h = np.empty(N, dtype=np.float64)
    if want_lists:
        d = np.full((N, k), np.inf); idx = np.full((N, k), N, dtype=np.int64)
        lib().orc_knn_brute(_p(pos), C.c_int64(N), C.c_int(k), C.c_double(box or 0.0), _p(h), _p(d), _p(idx))
        return h, d, idx
    lib().orc_knn_brute(_p(pos), C.c_int64(N), C.c_int(k), C.c_double(box or 0.0), _p(h), None, None)
    return h


def knn_scipy(pos, k, box=None, workers=1):
    """The reference's actual k-NN arithmetic: scipy.spatial.KDTree (io/SWIFT/_SnapshotSWIFT.py:69-82)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(pos, boxsize=box if box else None)
    d, i = tree.query(pos, k=k, workers=workers)
    if k == 1:
        d = d[:, None]; i = i[:, None]
    return d[:, k - 1].copy(), d, i


def max_threads():
    return int(lib().orc_max_threads())


# ---- the reference's own compiled path (oracle/_ref) ------------------------------------------------
def reference_available():
    suffix = __import__("sysconfig").get_config_var("EXT_SUFFIX")
    return os.path.exists(os.path.join(_HERE, "_ref", "pkg", "astro_sph_tools", "tools", "projections", "_kernels" + suffix))


def reference_module():
    """Import the reference's own create_image / quartic_spline_kernel from oracle/_ref (built by
    build_ref.py from /root/reference with non-arithmetic patches only)."""
    p = os.path.join(_HERE, "_ref", "pkg")
    if not reference_available():
        from . import build_ref
        if not build_ref.build():
            raise RuntimeError("oracle/_ref is not built and /root/reference is not present")
    if p not in sys.path:
        sys.path.insert(0, p)
    import importlib
    mod = importlib.import_module("astro_sph_tools.tools.projections")
    axes = importlib.import_module("astro_sph_tools._CoordinateAxes")
    return mod, axes.CoordinateAxes


# ---- N-D linear table interpolation ----------------------------------------------------------------
def table_interp(table, axes, x, fill_value=-np.inf):
    """numpy restatement of what IonisationTableBase.__call__ computes (reference data_structures/_IonisationTable.py:44-52:
    scipy.interpolate.RegularGridInterpolator(axes, table, bounds_error=False, fill_value=-inf)(x)).  The arithmetic lives
    in scipy (third party, un-pinned by the reference, 1.18.1 in this image): interval search of _rgi_cython.find_indices,
    corner sum of _rgi.py:_evaluate_linear in itertools.product order, then fill and NaN overrides of __call__.
    Pinned bit-for-bit against scipy itself in tests/test_table_oracle_cpu.py.  x: (N, ndim)."""
    import itertools
    table = np.asarray(table, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).reshape(-1, table.ndim)
    idx, y = [], []
    nan = np.any(np.isnan(x), axis=1)
    oob = np.zeros(x.shape[0], dtype=bool)
    for d, g in enumerate(axes):
        g = np.asarray(g, dtype=np.float64)
        xd = x[:, d]
        oob |= (xd < g[0]) | (xd > g[-1])
        i = np.clip(np.searchsorted(g, xd, side="right") - 1, 0, len(g) - 2)      # x[i] <= xval < x[i+1], edges clamped
        with np.errstate(invalid="ignore"):
            y.append((xd - g[i]) / (g[i + 1] - g[i]))
        idx.append(i)
    value = np.zeros(x.shape[0])
    with np.errstate(invalid="ignore"):
        if table.ndim == 2:
            # scipy's compiled 2-D fast path (_rgi_cython.evaluate_linear_2d, taken by _rgi.py __call__ for 2-D float64
            # tables) associates differently: (v * w0) * w1, summed left to right without the leading zero
            (i0, i1), (y0, y1) = idx, y
            value = (table[i0, i1] * (1 - y0) * (1 - y1) + table[i0, i1 + 1] * (1 - y0) * y1
                     + table[i0 + 1, i1] * y0 * (1 - y1) + table[i0 + 1, i1 + 1] * y0 * y1)
        else:
            for corner in itertools.product((0, 1), repeat=table.ndim):
                w = np.ones(x.shape[0])
                for d, up in enumerate(corner):
                    w = w * (y[d] if up else 1 - y[d])
                value = value + table[tuple(idx[d] + up for d, up in enumerate(corner))] * w
    value = np.asarray(value, dtype=np.float64)
    value[oob] = fill_value
    value[nan] = np.nan
    return value


def table_interp_scipy(table, axes, x, fill_value=-np.inf):
    """the call the reference makes (_IonisationTable.py:44-52)"""
    from scipy.interpolate import RegularGridInterpolator
    return RegularGridInterpolator(tuple(axes), table, bounds_error=False, fill_value=fill_value)(x)
