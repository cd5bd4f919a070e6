"""TEST INFRASTRUCTURE: generates tests/golden/reorder/*.npz by running the REFERENCE's own ArrayReorder and ArrayReorder_2
(/root/reference/src/astro_sph_tools/tools/_ArrayReorder.py:815-1038 and :659-812) on seeded ID sets.  The file is executed
from where it lies; its imports of packages that are absent here (`unyt`, `QuasarCode`, `QuasarCode.MPI`: units and MPI
plumbing, no matching arithmetic) are satisfied by stub modules.  Run in the build container only:
    python oracle/gen_golden_reorder.py"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/astro_sph_tools/tools/_ArrayReorder.py"


def load_reference_module():
    unyt = types.ModuleType("unyt")

    class unyt_array(np.ndarray):          # never instantiated by the cases below (plain numpy inputs)
        pass

    class unyt_quantity(unyt_array):
        pass

    unyt.unyt_array, unyt.unyt_quantity = unyt_array, unyt_quantity
    qc = types.ModuleType("QuasarCode"); qc.__path__ = []

    class Console:
        @staticmethod
        def print_debug(*a, **k): pass
        print_verbose_warning = print_verbose_info = print_info = print_warning = print_debug

    qc.Console = Console
    mpi = types.ModuleType("QuasarCode.MPI")

    class MPI_Config:
        comm = None; comm_size = 1; rank = 0; root = 0; is_root = True

    mpi.MPI_Config = MPI_Config
    mpi.mpi_barrier = lambda *a, **k: None
    mpi.synchronyse = lambda *a, **k: None
    mpi.mpi_gather_array = lambda a, *x, **k: a
    mpi.mpi_scatter_array = lambda a, *x, **k: a
    for name, m in (("unyt", unyt), ("QuasarCode", qc), ("QuasarCode.MPI", mpi)):
        sys.modules.setdefault(name, m)
    spec = importlib.util.spec_from_file_location("ref_array_reorder", SRC)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def cases():
    """(name, source_ids, target_ids, source_filter, target_filter, data, default_value)"""
    out = []
    rng = np.random.default_rng(20261018)
    # 1. the use at io/EAGLE/_CatalogueSUBFIND.py:292-295: snapshot IDs -> wanted IDs, group numbers with a default
    pool = rng.permutation(np.arange(10 ** 9, 10 ** 9 + 30000, dtype=np.int64) * 7919)
    src = pool[:9000].copy()
    tgt = rng.permutation(np.concatenate([pool[:9000][:3000], pool[9000:13000]]))
    out.append(("subfind_like", src, tgt, None, None, rng.integers(0, 1 << 30, 9000).astype(np.int64), np.int64(1 << 30)))
    # 2. a permutation (lossless), 2-D payload
    src = rng.permutation(5000).astype(np.int64)
    out.append(("permutation_2d", src, rng.permutation(src), None, None, rng.normal(size=(5000, 3)), None))
    # 3. filters on both sides
    src = rng.permutation(8000).astype(np.int64); tgt = rng.permutation(8000)[:5000].astype(np.int64)
    out.append(("filters", src, tgt, rng.random(8000) < 0.6, rng.random(5000) < 0.7, rng.normal(size=8000), -1.0))
    # 4. disjoint sets (no match), tiny source
    out.append(("disjoint", np.array([5, 3, 9], dtype=np.int64), np.arange(100, 140, dtype=np.int64), None, None,
                np.array([1.0, 2.0, 3.0]), 0.5))
    # 5. target is a subset of the source in another order (result exact, reduction)
    src = rng.permutation(6000).astype(np.int64)
    out.append(("subset", src, rng.permutation(src)[:2500], None, None, rng.integers(-50, 50, 6000).astype(np.int32), None))
    return out


if __name__ == "__main__":
    mod = load_reference_module()
    outdir = os.path.join(ROOT, "tests", "golden", "reorder")
    os.makedirs(outdir, exist_ok=True)
    for name, src, tgt, sf, tf, data, default in cases():
        r = mod.ArrayReorder.create(src, tgt, sf, tf)
        kw = {} if default is None else dict(default_value=default)
        res = r(data, **kw)
        back_default = np.array(-12345, dtype=res.dtype) if res.dtype.kind in "iu" else np.array(np.nan, dtype=res.dtype)
        back = r.reverse(res, default_value=back_default)
        save = dict(source_ids=src, target_ids=tgt, data=data, result=res, source_filter=r.source_filter, target_filter=r.target_filter,
                    matched=np.int64(r.matched_items), reverse_result=back, reverse_default=back_default,
                    flags=np.array([r.uses_all_inputs, r.all_outputs_matched, r.lossless, r.matches_are_reduction,
                                    r.results_are_expansion, r.results_are_subset, r.results_are_superset]))
        if sf is not None: save["source_order_filter"] = sf
        if tf is not None: save["target_order_filter"] = tf
        if default is not None: save["default_value"] = np.asarray(default)
        # the second single-process class of the reference (np.intersect1d based, :760-812) agrees with the first wherever it
        # runs at all: its call (:713) assigns through the caller's target FILTER, not the matched mask, so it raises as soon as
        # a target element has no match -- checked here for the cases in which every target is matched
        if r.all_outputs_matched:
            r2 = mod.ArrayReorder_2.create(src, tgt, sf, tf)
            assert np.array_equal(r2(data, **kw), res), name
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **save)
        print("wrote", path, os.path.getsize(path), "bytes; matched", r.matched_items, "of", len(src), "->", len(tgt))
