"""Pins the ID-matching oracle (tests/reorder_util.py) to results of the reference's OWN ArrayReorder class, run in the build
container with unyt / QuasarCode stubs (oracle/gen_golden_reorder.py -> tests/golden/reorder/*.npz)."""
import numpy as np
import pytest

from reorder_util import golden_reorder_cases, load_reorder, reference_reorder


@pytest.mark.parametrize("name", golden_reorder_cases())
def test_restatement_equals_reference_class(name):
    g = load_reorder(name)
    out, fwd, bwd = reference_reorder(g["source_ids"], g["target_ids"], g["data"], g["default_value"], g["source_order_filter"],
                                      g["target_order_filter"])
    assert np.array_equal(fwd, g["source_filter"]) and np.array_equal(bwd, g["target_filter"]) and bwd.sum() == g["matched"]
    if g["default_value"] is None:
        assert np.array_equal(out[bwd], g["result"][bwd])
    else:
        assert np.array_equal(out, g["result"])
    back, _, _ = reference_reorder(g["target_ids"], g["source_ids"], g["result"], g["reverse_default"][()], g["target_order_filter"],
                                   g["source_order_filter"])
    assert np.array_equal(back, g["reverse_result"], equal_nan=True)


def test_there_are_golden_cases():
    assert len(golden_reorder_cases()) >= 5
