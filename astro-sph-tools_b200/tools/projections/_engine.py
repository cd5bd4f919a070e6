"""Device-side driver of the 2-D projection: owns the torch buffers (inputs, maps, workspace) and calls the
C ABI (include/astro_sph_b200.h: ast_project2d).  PyTorch is used for device memory and streams only.

``Projector2D`` keeps its workspace between calls, so repeated projections of same-sized inputs (bench loops,
many maps of one snapshot) do not reallocate.
"""
import ctypes as C
import threading

import numpy as np

from ... import _lib
from ..._CoordinateAxes import CoordinateAxes


def _axis_index(projection_axis):
    if isinstance(projection_axis, CoordinateAxes):
        return projection_axis.value
    if hasattr(projection_axis, "value") and projection_axis.value in (0, 1, 2):      # the reference's own enum
        return int(projection_axis.value)
    if isinstance(projection_axis, str):
        return CoordinateAxes.from_string(projection_axis).value
    if isinstance(projection_axis, (int, np.integer)) and 0 <= int(projection_axis) <= 2:
        return int(projection_axis)
    raise ValueError(f"projection_axis must be a CoordinateAxes member, got {projection_axis!r}")


def batch_cuts(n, nb, ramp=True):
    """Boundaries of the host batches of project_host: nb batches of bn = ceil(n / nb) particles, preceded (ramp) by a
    quarter and a half batch so that the first, un-hidden host-to-device copy is short and every later copy is shorter
    than the deposition it hides behind.  Returns (bn, cuts) with cuts[0] = 0 and cuts[-1] = n."""
    bn = -(-int(n) // max(int(nb), 1))
    cuts = [0]
    if ramp and nb >= 2:
        for frac in (4, 2):
            cuts.append(min(n, cuts[-1] + max(1, bn // frac)))
    while cuts[-1] < n:
        cuts.append(min(n, cuts[-1] + bn))
    return bn, cuts


class Projector2D:
    """Reusable projection context on one CUDA device."""

    def __init__(self, device=None, pair_capacity=None, huge_capacity=1 << 20, small_max_px=-1, huge_min_tiles=-1):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        torch = self.torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.pair_capacity = pair_capacity
        self.huge_capacity = int(huge_capacity)
        self.small_max_px = int(small_max_px)
        self.huge_min_tiles = int(huge_min_tiles)
        self._ws = None
        self.last_stats = None
        self._lock = threading.RLock()        # one workspace and one set of staging buffers: calls on this engine serialise

    # ---- parameters -------------------------------------------------------------------------------------
    def _set_kernel(self, p, kernel):
        """built-in kernel by name / id, or a TabulatedKernel (arbitrary kernel_func callable)"""
        if isinstance(kernel, str):
            p.kernel_id = _lib.KERNEL_IDS[kernel]
        elif hasattr(kernel, "device_table"):
            tab = kernel.device_table(self.torch, self.device)
            self._kernel_table = tab                               # keep the device table alive for the call
            p.kernel_id = _lib.KERNEL_TABLE
            p.kernel_table = tab.data_ptr()
            p.kernel_table_n = tab.shape[0]
            p.kernel_dim = kernel.dim
        else:
            p.kernel_id = int(kernel)

    @staticmethod
    def _order_flags(presort):
        """flags for the optional spatial pre-ordering (presort.cuh): 'auto' lets the library look at a sample of the input order
        on every call (one small kernel and one stream synchronisation, ~20 us; sets below 65 536 particles are never touched),
        'always' orders without looking, 'never' leaves the input alone.  Both paths give the same map up to the order of the
        float additions."""
        if presort in (None, False, "never"):
            return 0
        if presort in (True, "always"):
            return _lib.FLAG_ORDER_ALWAYS
        if presort in ("auto", "sample"):
            return _lib.FLAG_ORDER_AUTO
        raise ValueError("presort must be 'auto', 'always' or 'never'")

    def _params(self, n, image_size, axis, bounds, kernel, n_prop, periodic, box, timing, accumulate):
        p = _lib.Project2DParams()
        p.n = int(n)
        p.axis = _axis_index(axis)
        p.nx, p.ny = int(image_size[0]), int(image_size[1])
        self._set_kernel(p, kernel)
        p.n_prop = int(n_prop)
        p.flags = (_lib.FLAG_PERIODIC if periodic else 0) | (_lib.FLAG_TIMING if timing else 0) | \
                  (_lib.FLAG_ACCUMULATE if accumulate else 0)
        p.x_min, p.x_max, p.y_min, p.y_max = (float(v) for v in bounds)
        if periodic:
            if box is None:
                raise ValueError("periodic=True needs box_size")
            if np.isscalar(box):
                ba = bb = float(box)
            else:
                box = tuple(float(v) for v in box)
                cols = CoordinateAxes(p.axis).plane_columns
                ba, bb = (box[cols[0]], box[cols[1]]) if len(box) == 3 else box
            p.box_a, p.box_b = ba, bb
        p.small_max_px = self.small_max_px
        p.huge_min_tiles = self.huge_min_tiles
        cap = self.pair_capacity
        if cap is None:
            cap = min(max(12 * int(n), 1 << 20), 1 << 30)
        p.pair_capacity = int(cap)
        p.huge_capacity = int(self.huge_capacity)
        return p

    def _workspace(self, p):
        need = C.c_size_t(0)
        _lib.check(self.lib.ast_project2d_workspace_bytes(C.byref(p), C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != self.device:
            self._ws = None
            self._ws = self.torch.empty(need.value, dtype=self.torch.uint8, device=self.device)
        return self._ws

    # ---- device-resident call -----------------------------------------------------------------------------
    def project(self, pos, h, props, image_size, axis, bounds, kernel="cubic_spline_3d", periodic=False, box=None,
                out=None, timing=False, accumulate=False, stream=None, presort="never"):
        """pos (N,3), h (N,), props = tensor (N,) or list of <= 2 tensors: float64 CUDA tensors.
        Returns the float64 CUDA map(s): (nx,ny) for a single tensor, (P,nx,ny) for a list.
        presort: 'never' (default for device-resident data: the caller knows its order, and the look costs one stream
        synchronisation, 0.06 ms -- 4 % of the step when every support is below a pixel), 'auto' (look at a sample of the
        input order and project a tile-ordered copy if it is incoherent; what the host paths use), 'always'."""
        torch = self.torch
        single = not isinstance(props, (list, tuple))
        plist = [props] if single else list(props)
        if not 1 <= len(plist) <= _lib.MAX_PROPS:
            raise ValueError(f"between 1 and {_lib.MAX_PROPS} weight arrays per pass, got {len(plist)}")
        n = pos.shape[0]
        for t, shape in [(pos, (n, 3)), (h, (n,))] + [(q, (n,)) for q in plist]:
            if t.dtype != torch.float64 or tuple(t.shape) != shape or not t.is_cuda or not t.is_contiguous():
                raise ValueError("device inputs must be contiguous float64 CUDA tensors of shapes (N,3), (N,), (N,)")
        p = self._params(n, image_size, axis, bounds, kernel, len(plist), periodic, box, timing, accumulate)
        p.flags |= self._order_flags(presort)
        if out is None:
            out = torch.empty((len(plist), p.nx, p.ny), dtype=torch.float64, device=self.device)
        elif out.dtype != torch.float64 or out.numel() != len(plist) * p.nx * p.ny or not out.is_contiguous():
            raise ValueError("out must be a contiguous float64 tensor of n_prop*nx*ny elements")
        prop_ptrs = (C.c_void_p * _lib.MAX_PROPS)(*[q.data_ptr() for q in plist] + [None] * (_lib.MAX_PROPS - len(plist)))
        stats = _lib.Project2DStats()
        with torch.cuda.device(self.device), self._lock:
            # pair_capacity and huge_capacity only size windows: the library walks any number of pairs / large-h entries
            # through them in rounds and never fails for lack of capacity once it has started depositing
            ws = self._workspace(p)
            _lib.check(self.lib.ast_project2d(C.byref(p), _lib.ptr(pos), _lib.ptr(h), prop_ptrs, _lib.ptr(out), _lib.ptr(ws),
                                              C.c_size_t(ws.numel()), _lib.stream_ptr(stream), C.byref(stats)))
        self.last_stats = dict(n_pairs=stats.n_pairs, n_huge=stats.n_huge, n_rounds=stats.n_rounds,
                               n_launches=stats.n_launches, stage_ms=list(stats.stage_ms), reordered=bool(stats.reordered))
        out = out.view(len(plist), p.nx, p.ny)
        return out[0] if single else out

    # ---- host call (numpy in, numpy out): what create_image uses -------------------------------------------
    def project_host(self, positions, smoothing_lengths, props, image_size, axis, bounds, kernel="cubic_spline_3d",
                     periodic=False, box=None, stream=None, batch_particles=1 << 22, return_device=False, ramp=True):
        """Host arrays in, host map out.  The particle arrays are streamed to the device in batches of
        `batch_particles` through two staging buffers on a copy stream, so the host-to-device transfer of batch b+1
        overlaps the deposition of batch b (deposition is linear in particles: batches accumulate into the same map).
        With ramp=True the first batches are a quarter and a half of `batch_particles`, so the first (unhidden) copy is short.
        Pinned host arrays make the copies truly asynchronous; pageable ones still overlap with the running kernels."""
        torch = self.torch
        single = not isinstance(props, (list, tuple))
        plist = [props] if single else list(props)
        dev = self.device
        n = int(positions.shape[0])
        nb = max(1, min(16, -(-n // int(batch_particles)))) if n > 0 else 1
        if nb == 1:
            # float32 host arrays (the ingestion shim's opt-in) cross PCIe as float32 and are widened on the device: exact
            to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True).to(torch.float64)
            with torch.cuda.device(dev), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
                pos_d, h_d = to_dev(positions), to_dev(smoothing_lengths)      # uploads ordered on the stream that computes
                props_d = [to_dev(q) for q in plist]
            out = self.project(pos_d, h_d, props_d[0] if single else props_d, image_size, axis, bounds, kernel, periodic, box,
                               stream=stream, presort="auto")
        else:
            positions = np.ascontiguousarray(positions)
            smoothing_lengths = np.ascontiguousarray(smoothing_lengths)
            plist = [np.ascontiguousarray(q) for q in plist]
            bn, cuts = batch_cuts(n, nb, ramp)
            nb = len(cuts) - 1
            with torch.cuda.device(dev), self._lock:
                compute = stream if stream is not None else torch.cuda.current_stream()
                if getattr(self, "_copy_stream", None) is None:
                    self._copy_stream = torch.cuda.Stream(device=dev)
                copy = self._copy_stream
                key = (bn, len(plist))
                if getattr(self, "_stage_key", None) != key:
                    self._stage = [dict(pos=torch.empty((bn, 3), dtype=torch.float64, device=dev),
                                        h=torch.empty(bn, dtype=torch.float64, device=dev),
                                        props=[torch.empty(bn, dtype=torch.float64, device=dev) for _ in plist]) for _ in range(2)]
                    self._stage_key = key
                ready = [torch.cuda.Event(), torch.cuda.Event()]
                free = [torch.cuda.Event(), torch.cuda.Event()]
                out = torch.empty((len(plist), int(image_size[0]), int(image_size[1])), dtype=torch.float64, device=dev)
                copy.wait_stream(compute)
                n_launch = 0
                for b in range(nb):
                    lo, hi = cuts[b], cuts[b + 1]
                    m = hi - lo
                    if m <= 0:
                        break
                    k = b & 1
                    st = self._stage[k]
                    with torch.cuda.stream(copy):
                        if b >= 2:
                            copy.wait_event(free[k])

                        def put(dst, src):
                            src = torch.from_numpy(src)
                            if src.dtype == torch.float64:
                                dst.copy_(src, non_blocking=True)
                            else:         # float32 over PCIe, widened on the device (a cross-dtype copy_ would convert on the CPU)
                                dst.copy_(src.to(dev, non_blocking=True))
                        put(st["pos"][:m], positions[lo:hi])
                        put(st["h"][:m], smoothing_lengths[lo:hi])
                        for dst, src in zip(st["props"], plist):
                            put(dst[:m], src[lo:hi])
                        ready[k].record(copy)
                    compute.wait_event(ready[k])
                    self.project(st["pos"][:m], st["h"][:m], [q[:m] for q in st["props"]], image_size, axis, bounds, kernel,
                                 periodic, box, out=out, accumulate=b > 0, stream=compute, presort="auto")
                    n_launch += self.last_stats["n_launches"]
                    free[k].record(compute)
                self.last_stats["n_launches"] = n_launch
                self.last_stats["n_batches"] = nb
            out = out[0] if single else out
        if return_device:
            return out
        # device -> pinned host block (from torch's caching host allocator, so no page faults and full PCIe rate); the
        # numpy array returned is a view that owns the block, i.e. a fresh array for the caller like the reference's
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        with torch.cuda.device(dev), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream()):
            host.copy_(out, non_blocking=True)                               # ordered behind the deposition on ITS stream
            torch.cuda.current_stream().synchronize()
        return host.numpy()
