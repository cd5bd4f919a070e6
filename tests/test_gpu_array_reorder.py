"""-m gpu: ArrayReorder / match_ids (GPU hash join) against a numpy restatement of the reference's ArrayReorder
(tools/_ArrayReorder.py:988-1038 create, :937-961 call).  Integer work: results must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


from reorder_util import golden_reorder_cases, load_reorder, reference_reorder


@pytest.mark.parametrize("name", golden_reorder_cases())
def test_reorder_equals_the_reference_class_results(name):
    """golden results of the reference's own ArrayReorder (oracle/gen_golden_reorder.py): forward call, filters, flags, reverse"""
    from astro_sph_tools_b200.tools import ArrayReorder
    g = load_reorder(name)
    r = ArrayReorder.create(g["source_ids"], g["target_ids"], g["source_order_filter"], g["target_order_filter"])
    assert np.array_equal(r.source_filter, g["source_filter"]) and np.array_equal(r.target_filter, g["target_filter"])
    assert r.matched_items == g["matched"]
    flags = [r.uses_all_inputs, r.all_outputs_matched, r.lossless, r.matches_are_reduction, r.results_are_expansion,
             r.results_are_subset, r.results_are_superset]
    assert np.array_equal(np.array(flags), g["flags"])
    kw = {} if g["default_value"] is None else dict(default_value=g["default_value"])
    out = r(g["data"], **kw)
    assert out.dtype == g["result"].dtype and out.shape == g["result"].shape
    assert np.array_equal(out[r.target_filter], g["result"][r.target_filter])
    if g["default_value"] is not None:
        assert np.array_equal(out, g["result"])
    back = r.reverse(g["result"], default_value=g["reverse_default"][()])
    assert np.array_equal(back, g["reverse_result"], equal_nan=True)


@pytest.mark.parametrize("n_src,n_tgt", [(1000, 1000), (50000, 20000), (3, 70000), (200000, 200000)])
def test_reorder_matches_reference_arithmetic(n_src, n_tgt):
    from astro_sph_tools_b200.tools import ArrayReorder
    rng = np.random.default_rng(n_src + n_tgt)
    pool = rng.permutation(np.arange(10 ** 6, 10 ** 6 + 3 * max(n_src, n_tgt), dtype=np.int64) * 7919)
    src = pool[:n_src].copy()
    tgt = rng.permutation(np.concatenate([pool[:n_src][: n_tgt // 2], pool[n_src:n_src + n_tgt]])[:n_tgt])
    data = rng.normal(size=(n_src, 3))
    ref, fwd, bwd = reference_reorder(src, tgt, data, -1.0)
    r = ArrayReorder.create(src, tgt)
    out = r(data, default_value=-1.0)
    assert np.array_equal(out, ref)
    assert np.array_equal(r.source_filter, fwd) and np.array_equal(r.target_filter, bwd)
    assert r.matched_items == bwd.sum() and r.input_length == n_src and r.output_length == n_tgt
    # the reverse object maps target-ordered data back
    back = r.reverse(out[:, 0].copy(), default_value=np.nan)
    assert np.array_equal(back[fwd], data[fwd, 0]) and np.isnan(back[~fwd]).all()


def test_filters_dtypes_and_errors():
    from astro_sph_tools_b200.tools import ArrayReorder, match_ids
    rng = np.random.default_rng(5)
    src = rng.permutation(5000).astype(np.uint64); tgt = rng.permutation(5000)[:3000].astype(np.int64)
    sf = rng.random(5000) < 0.6; tf = rng.random(3000) < 0.7
    data = rng.integers(0, 1 << 40, 5000)
    ref, fwd, bwd = reference_reorder(src.astype(np.int64), tgt, data, -7, sf, tf)
    r = ArrayReorder.create(src, tgt, sf, tf)
    assert np.array_equal(r(data, default_value=-7), ref)
    idx = match_ids(src, tgt)
    assert np.array_equal(src[idx].astype(np.int64), tgt)                  # every target present in the source
    full = ArrayReorder.create(src, src[::-1].copy())
    assert full.lossless and np.array_equal(full(data), data[::-1])
    with pytest.raises(ValueError, match="no default value"):
        r(data)
    empty = ArrayReorder.create(np.zeros(0, dtype=np.int64), tgt)
    assert empty.matched_items == 0 and np.array_equal(empty(np.zeros(0), default_value=3.0), np.full(3000, 3.0))
