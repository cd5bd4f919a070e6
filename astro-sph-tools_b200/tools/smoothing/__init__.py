"""Smoothing lengths from k nearest neighbours: replaces the KDTree branch of the reference's
``SnapshotSWIFT.get_smoothing_lengths`` (io/SWIFT/_SnapshotSWIFT.py:58-85):

    tree = KDTree(positions); h = tree.query(positions, k=32)[0][:, 31]

i.e. h_i = distance to the K-th nearest particle, the particle itself counted as the first, Euclidean,
non-periodic, K = 32 hard-coded (``N_NABOURS``, :63, with a TODO to make it a setting -- here it is ``k=``).
``box_size`` adds scipy's ``boxsize`` semantics (the reference uses it in _scripts/find_nearest_haloes.py:207-210).
Distances are bit-equal to scipy's float64 arithmetic.  CUDA only (ast_knn_h); no CPU fallback.
"""
import ctypes as C
import threading

import numpy as np

from ... import _lib

DEFAULT_K = 32          # N_NABOURS, io/SWIFT/_SnapshotSWIFT.py:63


class SmoothingLengthSolver:
    """Reusable k-NN context on one CUDA device (keeps its workspace)."""

    def __init__(self, device=None, cell_target=0.0):
        self.torch = _lib.require_cuda()
        self.lib = _lib.load()
        torch = self.torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cell_target = float(cell_target)
        self._ws = None
        self._lock = threading.RLock()        # one workspace: calls on this solver serialise

    def solve(self, pos, k=DEFAULT_K, box_size=None, q_begin=0, q_count=0, want_neighbours=False, want_distances=False,
              stream=None, kernel="select", full_build=False, cell_target=None):
        """pos: (N,3) float64 CUDA tensor.  Returns h (Q,) [, idx (Q,k) int32] [, dist (Q,k)] as CUDA tensors where
        Q = q_count or N.  Multi-GPU use: every rank passes all positions and its own [q_begin, q_begin+q_count)."""
        torch = self.torch
        if pos.dtype != torch.float64 or pos.ndim != 2 or pos.shape[1] != 3 or not pos.is_cuda or not pos.is_contiguous():
            raise ValueError("pos must be a contiguous float64 CUDA tensor of shape (N, 3)")
        n = pos.shape[0]
        p = _lib.KnnParams()
        p.n = n; p.k = int(k); p.flags = {"select": 0, "lockstep": 8, "diverging": 1}[kernel] | (4 if full_build else 0)      # AST_KNN_* query-kernel selection (csrc/knn.cu): select = selection over blocks of cells + lock-step for what it cannot verify (h only; neighbour lists always take the lock-step kernel)
        p.box = float(box_size) if box_size else 0.0
        if p.box <= 0.0 and n > 0:
            lo, hi = (t.cpu() for t in torch.aminmax(pos, dim=0))            # one pass, one read-back
            if not (torch.isfinite(lo).all() and torch.isfinite(hi).all()):
                raise ValueError("positions must be finite")
            for c in range(3):
                p.lo[c] = float(lo[c]); p.hi[c] = float(hi[c])
        elif n > 0:
            mn, mx = (float(v) for v in torch.stack(torch.aminmax(pos)).cpu())     # one pass over the positions, one read-back
            if not (mn >= 0.0 and mx < p.box):
                raise ValueError("periodic k-NN needs 0 <= x < box_size (scipy boxsize semantics)")
        # mean particles per cell if the set filled the whole box; a caller whose set fills a fraction f of it (a slab of a
        # multi-GPU decomposition) passes 2 f to keep ~2 particles per OCCUPIED cell
        p.cell_target = float(cell_target) if cell_target else self.cell_target
        p.q_begin = int(q_begin); p.q_count = int(q_count)
        nq = int(q_count) if q_count and q_count > 0 else n
        need = C.c_size_t(0)
        _lib.check(self.lib.ast_knn_workspace_bytes(C.byref(p), C.byref(need)))
        h = torch.empty(nq, dtype=torch.float64, device=self.device)
        idx = torch.empty((nq, k), dtype=torch.int32, device=self.device) if want_neighbours else None
        dist = torch.empty((nq, k), dtype=torch.float64, device=self.device) if want_distances else None
        with torch.cuda.device(self.device), self._lock:
            if self._ws is None or self._ws.numel() < need.value:
                self._ws = None
                self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ast_knn_h(C.byref(p), _lib.ptr(pos), _lib.ptr(h), _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(self._ws),
                                          C.c_size_t(self._ws.numel()), _lib.stream_ptr(stream)))
        out = (h,)
        if want_neighbours:
            out += (idx,)
        if want_distances:
            out += (dist,)
        return out if len(out) > 1 else h


    def _knn_params(self, data, k, box_size):
        torch = self.torch
        if data.dtype != torch.float64 or data.ndim != 2 or data.shape[1] != 3 or not data.is_cuda or not data.is_contiguous():
            raise ValueError("positions must be a contiguous float64 CUDA tensor of shape (N, 3)")
        p = _lib.KnnParams()
        p.n = data.shape[0]; p.k = int(k); p.flags = 0
        p.box = float(box_size) if box_size else 0.0
        if p.box <= 0.0:
            lo = data.min(dim=0).values.cpu(); hi = data.max(dim=0).values.cpu()
            if not (torch.isfinite(lo).all() and torch.isfinite(hi).all()):
                raise ValueError("positions must be finite")
            for c in range(3):
                p.lo[c] = float(lo[c]); p.hi[c] = float(hi[c])
        elif not (float(data.min()) >= 0.0 and float(data.max()) < p.box):
            raise ValueError("periodic k-NN needs 0 <= x < box_size (scipy boxsize semantics)")
        p.cell_target = self.cell_target
        return p

    def query(self, data, queries, k=1, box_size=None, stream=None):
        """k nearest DATA points of each QUERY point: (dist (M,k) float64, idx (M,k) int32) CUDA tensors, ascending."""
        torch = self.torch
        if queries.dtype != torch.float64 or queries.ndim != 2 or queries.shape[1] != 3 or not queries.is_cuda or not queries.is_contiguous():
            raise ValueError("queries must be a contiguous float64 CUDA tensor of shape (M, 3)")
        if data.shape[0] == 0:
            raise ValueError("empty data set")
        p = self._knn_params(data, k, box_size)
        if p.box > 0.0 and queries.shape[0] and not (float(queries.min()) >= 0.0 and float(queries.max()) < p.box):
            raise ValueError("periodic queries must satisfy 0 <= x < box_size")
        need = C.c_size_t(0)
        _lib.check(self.lib.ast_knn_workspace_bytes(C.byref(p), C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        m = queries.shape[0]
        dist = torch.empty((m, k), dtype=torch.float64, device=self.device)
        idx = torch.empty((m, k), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device), self._lock:
            _lib.check(self.lib.ast_knn_query(C.byref(p), _lib.ptr(data), _lib.ptr(queries), C.c_int64(m), _lib.ptr(dist), _lib.ptr(idx),
                                              _lib.ptr(self._ws), C.c_size_t(self._ws.numel()), _lib.stream_ptr(stream)))
        return dist, idx


_default = {}


def _solver(device=None):
    torch = _lib.require_cuda()
    key = torch.cuda.current_device() if device is None else torch.device(device).index
    if key not in _default:
        _default[key] = SmoothingLengthSolver(device)
    return _default[key]


def compute_smoothing_lengths_device(pos, k=DEFAULT_K, box_size=None, **kw):
    return _solver(pos.device).solve(pos, k, box_size, **kw)


def compute_smoothing_lengths(positions, k=DEFAULT_K, box_size=None, return_neighbours=False):
    """numpy in / numpy out.  positions (N,3) float64 -> h (N,) float64  [, neighbours (N,k) int32]."""
    positions = np.asarray(positions.value if hasattr(positions, "value") and not isinstance(positions, np.ndarray) else positions)
    if positions.dtype != np.float64:
        raise ValueError(f"Buffer dtype mismatch, expected 'double' but got '{positions.dtype.name}'")
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (N, 3)")
    torch = _lib.require_cuda()
    sol = _solver()
    pos_d = torch.from_numpy(np.ascontiguousarray(positions)).to(sol.device)
    res = sol.solve(pos_d, k, box_size, want_neighbours=return_neighbours)
    if return_neighbours:
        return res[0].cpu().numpy(), res[1].cpu().numpy()
    return res.cpu().numpy()


def get_smoothing_lengths(positions, n_neighbours=DEFAULT_K):
    """The reference's semantics exactly (K = 32, self included, non-periodic): io/SWIFT/_SnapshotSWIFT.py:62-83."""
    return compute_smoothing_lengths(positions, k=n_neighbours, box_size=None)


def nearest_neighbours(data_positions, query_positions, k=1, box_size=None):
    """``scipy.spatial.KDTree(data_positions, boxsize=box_size).query(query_positions, k)`` on the GPU: the nearest-halo
    lookup of the reference's CLI (_scripts/find_nearest_haloes.py:207-215: tree over halo centres, queried with particle
    positions).  numpy in / numpy out; returns (distances, indices), 1-D for k == 1 like scipy, else (M, k)."""
    torch = _lib.require_cuda()
    sol = _solver()
    d = torch.from_numpy(np.ascontiguousarray(data_positions, dtype=np.float64)).to(sol.device)
    q = torch.from_numpy(np.ascontiguousarray(query_positions, dtype=np.float64)).to(sol.device)
    dist, idx = sol.query(d, q, k, box_size)
    dist, idx = dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    return (dist[:, 0], idx[:, 0]) if k == 1 else (dist, idx)
