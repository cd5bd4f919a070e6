"""ArrayReorder: reorder per-particle arrays from one ID order to another.

Mirror of the reference's ``ArrayReorder`` (tools/_ArrayReorder.py:815-1038; used as
``ArrayReorder.create(snapshot_ids, wanted_ids)(group_numbers, default_value=...)`` in
io/EAGLE/_CatalogueSUBFIND.py:292-295): same ``create`` arguments, properties, ``reverse`` and call semantics
(``output[target_filter] = source_data[source_filter][order]``, unmatched outputs take ``default_value``).  The reference
matches IDs with argsort + np.isin / np.intersect1d on one CPU core; here the match is a GPU hash join (ast_match_ids)
and the data movement a GPU row gather (ast_gather_rows).  IDs must be unique within each (filtered) array, as the
reference assumes (``assume_unique=True``).  No CPU fallback.
"""
import ctypes as C
from typing import Any, Union

import numpy as np

from .. import _lib


def _match(source_ids, target_ids, source_filter, target_filter):
    """index of the matching source element for every target element (-1: none), on the GPU"""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    to_ids = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int64, copy=False).view(np.int64)).to(dev)
    to_mask = lambda m: None if m is None else torch.from_numpy(np.ascontiguousarray(m, dtype=np.uint8)).to(dev)
    s, t = to_ids(source_ids), to_ids(target_ids)
    sf, tf = to_mask(source_filter), to_mask(target_filter)
    need = C.c_size_t(0)
    _lib.check(lib.ast_match_ids_workspace_bytes(C.c_int64(s.numel()), C.byref(need)))
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    out = torch.empty(t.numel(), dtype=torch.int64, device=dev)
    _lib.check(lib.ast_match_ids(_lib.ptr(s), C.c_int64(s.numel()), _lib.ptr(sf), _lib.ptr(t), C.c_int64(t.numel()), _lib.ptr(tf),
                                 _lib.ptr(out), _lib.ptr(ws), C.c_size_t(need.value), _lib.stream_ptr()))
    return out.cpu().numpy()


class ArrayReorder:
    """Callable that rearranges the elements of an array from the source ID order into the target ID order.
    Use ``ArrayReorder.create`` and the ``reverse`` attribute (like the reference, not the constructor)."""

    def __init__(self, source_index_of_target: np.ndarray, input_length: int) -> None:
        self._t2s = source_index_of_target
        self._destination_filter = source_index_of_target >= 0
        self._source_filter = np.zeros(input_length, dtype=bool)
        self._source_filter[source_index_of_target[self._destination_filter]] = True
        self._n_matched = int(self._destination_filter.sum())
        self._reverse = None

    # ---- the reference's read-only properties (tools/_ArrayReorder.py:845-935) -------------------------------------
    reverse = property(lambda self: self._reverse)
    source_filter = property(lambda self: self._source_filter)
    target_filter = property(lambda self: self._destination_filter)
    input_length = property(lambda self: self._source_filter.shape[0])
    output_length = property(lambda self: self._destination_filter.shape[0])
    matched_items = property(lambda self: self._n_matched)
    uses_all_inputs = property(lambda self: self.input_length == self._n_matched)
    all_outputs_matched = property(lambda self: self.output_length == self._n_matched)
    lossless = property(lambda self: self.uses_all_inputs and self.all_outputs_matched)
    matches_are_reduction = property(lambda self: self.input_length > self._n_matched)
    results_are_expansion = property(lambda self: self.output_length > self._n_matched)
    results_are_subset = property(lambda self: self.matches_are_reduction and self.all_outputs_matched)
    results_are_superset = property(lambda self: self.results_are_expansion and self.uses_all_inputs)

    def __len__(self) -> int:
        return self.input_length

    def __call__(self, source_data: np.ndarray, /, output_array: Union[np.ndarray, None] = None,
                 default_value: Union[Any, None] = None) -> np.ndarray:
        """Reorder data (reference: tools/_ArrayReorder.py:937-961)."""
        units = None
        if hasattr(source_data, "units") and hasattr(source_data, "value"):          # unyt_array without importing unyt
            units, source_data = source_data.units, np.asarray(source_data.value)
        if not self.all_outputs_matched and output_array is None and default_value is None:
            raise ValueError("More output elements expected than matches but no default value provided and no output target "
                             "array to write matches to.")
        source_data = np.ascontiguousarray(source_data)
        if source_data.shape[0] != self.input_length:
            raise ValueError(f"source_data has {source_data.shape[0]} rows, expected {self.input_length}")
        if output_array is None:
            output_array = np.empty(shape=(self.output_length, *source_data.shape[1:]), dtype=source_data.dtype)
        if default_value is not None:
            output_array[~self._destination_filter] = default_value
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        row_bytes = source_data.dtype.itemsize * int(np.prod(source_data.shape[1:], dtype=np.int64))
        src = torch.from_numpy(source_data.view(np.uint8).reshape(-1)).to(dev)
        out_host = np.ascontiguousarray(output_array)
        if out_host.dtype != source_data.dtype or out_host.shape != (self.output_length, *source_data.shape[1:]):
            raise ValueError("output_array must have the dtype of source_data and shape (output_length, ...)")
        out = torch.from_numpy(out_host.view(np.uint8).reshape(-1).copy()).to(dev)
        idx = torch.from_numpy(self._t2s).to(dev)
        _lib.check(lib.ast_gather_rows(_lib.ptr(src), C.c_int64(row_bytes), _lib.ptr(idx), C.c_int64(self.output_length), _lib.ptr(out),
                                       _lib.stream_ptr()))
        result = out.cpu().numpy().view(source_data.dtype).reshape(out_host.shape)
        output_array[...] = result
        if units is not None:
            import unyt
            return unyt.unyt_array(output_array, units)
        return output_array

    @staticmethod
    def create(source_order: np.ndarray, target_order: np.ndarray, source_order_filter: Union[np.ndarray, None] = None,
               target_order_filter: Union[np.ndarray, None] = None) -> "ArrayReorder":
        """source_order / target_order: integer ID arrays.  The optional boolean filters restrict which elements may match
        without changing the input or output shapes (reference: tools/_ArrayReorder.py:964-1038)."""
        source_order = np.asarray(source_order); target_order = np.asarray(target_order)
        if source_order.ndim != 1 or target_order.ndim != 1 or source_order.dtype.kind not in "iu" or target_order.dtype.kind not in "iu":
            raise ValueError("source_order and target_order must be 1-D integer arrays")
        t2s = _match(source_order, target_order, source_order_filter, target_order_filter)
        s2t = _match(target_order, source_order, target_order_filter, source_order_filter)
        forwards = ArrayReorder(t2s, source_order.shape[0])
        backwards = ArrayReorder(s2t, target_order.shape[0])
        forwards._reverse = backwards
        backwards._reverse = forwards
        return forwards


def gather_rows_device(rows: np.ndarray, index: np.ndarray, default_value=None) -> np.ndarray:
    """rows[index] on the GPU (ast_gather_rows); rows of the result whose index is negative take default_value"""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    rows = np.ascontiguousarray(rows)
    out_host = np.empty((index.shape[0],) + rows.shape[1:], dtype=rows.dtype)
    if default_value is not None:
        out_host[...] = default_value
    row_bytes = rows.dtype.itemsize * int(np.prod(rows.shape[1:], dtype=np.int64))
    src = torch.from_numpy(rows.view(np.uint8).reshape(-1)).to(dev)
    out = torch.from_numpy(out_host.view(np.uint8).reshape(-1)).to(dev)
    idx = torch.from_numpy(np.ascontiguousarray(index, dtype=np.int64)).to(dev)
    _lib.check(lib.ast_gather_rows(_lib.ptr(src), C.c_int64(row_bytes), _lib.ptr(idx), C.c_int64(index.shape[0]), _lib.ptr(out),
                                   _lib.stream_ptr()))
    return out.cpu().numpy().view(rows.dtype).reshape(out_host.shape)


def match_ids(source_ids, target_ids, source_filter=None, target_filter=None) -> np.ndarray:
    """For every target ID the index of the equal source ID, -1 where there is none (GPU hash join)."""
    return _match(np.asarray(source_ids), np.asarray(target_ids), source_filter, target_filter)
