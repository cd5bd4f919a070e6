"""Thin test-side callers of the C ABI (through ctypes + torch buffers), used by the -m gpu tests."""
import ctypes as C

import numpy as np
import torch

from astro_sph_tools_b200 import _lib
from astro_sph_tools_b200.tools.projections import Projector2D


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_params(n, image_size, axis, bounds, kernel="cubic_spline_3d", n_prop=1, periodic=False, box=None,
                small_max_px=-1, huge_min_tiles=-1, pair_capacity=1 << 20, huge_capacity=1 << 16):
    p = _lib.Project2DParams()
    p.n = n; p.axis = axis; p.nx, p.ny = image_size; p.kernel_id = _lib.KERNEL_IDS[kernel]; p.n_prop = n_prop
    p.flags = _lib.FLAG_PERIODIC if periodic else 0
    p.x_min, p.x_max, p.y_min, p.y_max = bounds
    if periodic:
        p.box_a, p.box_b = box
    p.small_max_px = small_max_px; p.huge_min_tiles = huge_min_tiles
    p.pair_capacity = pair_capacity; p.huge_capacity = huge_capacity
    return p


def gpu_bin2d(pos, h, image_size, axis, bounds, periodic=False, box=None, small_max_px=16, huge_min_tiles=256,
              pair_capacity=1 << 22, huge_capacity=1 << 18):
    lib = _lib.load()
    n = len(h)
    n_img = 9 if periodic else 1
    p = make_params(n, image_size, axis, bounds, periodic=periodic, box=box, small_max_px=small_max_px,
                    huge_min_tiles=huge_min_tiles, pair_capacity=pair_capacity, huge_capacity=huge_capacity)
    need = C.c_size_t(0)
    _lib.check(lib.ast_project2d_workspace_bytes(C.byref(p), C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    pos_d, h_d = dev(pos), dev(h)
    bbox = torch.empty((n_img * n, 4), dtype=torch.int32, device="cuda")
    cls = torch.empty(n_img * n, dtype=torch.uint8, device="cuda")
    pe = torch.zeros(pair_capacity, dtype=torch.int64, device="cuda")
    ps = torch.zeros(pair_capacity, dtype=torch.int64, device="cuda")
    hg = torch.zeros(huge_capacity, dtype=torch.int64, device="cuda")
    counts = (C.c_int64 * 2)()
    _lib.check(lib.ast_bin2d(C.byref(p), _lib.ptr(pos_d), _lib.ptr(h_d), _lib.ptr(bbox), _lib.ptr(cls), _lib.ptr(pe), _lib.ptr(ps),
                             _lib.ptr(hg), counts, _lib.ptr(ws), C.c_size_t(need.value), _lib.stream_ptr()))
    torch.cuda.synchronize()
    npairs, nh = counts[0], counts[1]
    return dict(bbox=bbox.cpu().numpy(), cls=cls.cpu().numpy().reshape(n_img, n),
                pairs=pe.cpu().numpy()[:npairs].view(np.uint64), sorted=ps.cpu().numpy()[:npairs].view(np.uint64),
                huge=hg.cpu().numpy()[:nh].view(np.uint64))


def gpu_contrib_count(pos, h, image_size, axis, bounds, periodic=False, box=None):
    lib = _lib.load()
    p = make_params(len(h), image_size, axis, bounds, periodic=periodic, box=box)
    cnt = torch.empty(image_size, dtype=torch.int32, device="cuda")
    pos_d, h_d = dev(pos), dev(h)          # keep the tensors alive until the kernel has run
    _lib.check(lib.ast_contrib_count2d(C.byref(p), _lib.ptr(pos_d), _lib.ptr(h_d), _lib.ptr(cnt), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return cnt.cpu().numpy()


def gpu_sort(elems, bit_lo, n_bits):
    lib = _lib.load()
    n = len(elems)
    a = torch.from_numpy(elems.view(np.int64).copy()).cuda()
    b = torch.empty_like(a)
    need = C.c_size_t(0)
    _lib.check(lib.ast_sort_workspace_bytes(C.c_int64(n), C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device="cuda")
    flag = C.c_int(0)
    _lib.check(lib.ast_radix_sort_u64(_lib.ptr(a), _lib.ptr(b), C.c_int64(n), C.c_int(bit_lo), C.c_int(n_bits), _lib.ptr(ws),
                                      C.c_size_t(need.value), _lib.stream_ptr(), C.byref(flag)))
    torch.cuda.synchronize()
    return (b if flag.value else a).cpu().numpy().view(np.uint64)


def gpu_project(pos, h, props, image_size, axis, bounds, kernel="cubic_spline_3d", periodic=False, box=None, **engine_kw):
    eng = Projector2D(**engine_kw)
    single = not isinstance(props, (list, tuple))
    pl = [props] if single else list(props)
    out = eng.project(dev(pos), dev(h), dev(pl[0]) if single else [dev(q) for q in pl], image_size, axis, bounds, kernel,
                      periodic, box)
    torch.cuda.synchronize()
    return out.cpu().numpy(), eng.last_stats
