"""CPU: the table-interpolation restatement (oracle.table_interp) is pinned bit-for-bit against
scipy.interpolate.RegularGridInterpolator, the arithmetic the reference calls (data_structures/_IonisationTable.py:44-52;
scipy is third-party and un-pinned by the reference, 1.18.1 here), and against golden values produced by the reference's own
IonisationTableBase class (tests/golden/tables/ion_table_golden.npz, generator oracle/gen_golden_table.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from table_util import gas_state, synthetic_table


@pytest.mark.parametrize("uniform", [True, False])
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_restatement_bit_equals_scipy(oracle, ndim, uniform):
    table, axes = synthetic_table(uniform=uniform)
    sl = (slice(None),) * ndim + (3,) * (3 - ndim)
    table, axes = np.ascontiguousarray(table[sl]), axes[:ndim]
    x = gas_state(11 + ndim, 5000, axes)
    ref = oracle.table_interp_scipy(table, axes, x)
    got = oracle.table_interp(table, axes, x)
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.isnan(ref[40]) and np.isnan(ref[41]) and np.isneginf(ref[44]) and np.isneginf(ref[45]) and np.isneginf(ref[46:]).any()
    assert np.isfinite(ref[42]) and np.isfinite(ref[43])          # both corners of the grid are inside


def test_restatement_matches_reference_class_golden(oracle):
    g = np.load(os.path.join(GOLDEN_DIR, "tables", "ion_table_golden.npz"))
    axes = [g["axis0"], g["axis1"], g["axis2"]]
    assert np.array_equal(oracle.table_interp(g["table"], axes, g["x"]), g["call"], equal_nan=True)
    x3 = np.insert(g["x2"], 2, float(g["redshift"]), axis=1)
    assert np.array_equal(oracle.table_interp(g["table"], axes, x3), g["at_redshift"], equal_nan=True)


def test_table_with_infinite_entries_propagates_like_scipy(oracle):
    table, axes = synthetic_table(shape=(9, 11, 5))
    table[2, 3, 1] = -np.inf                       # log10 of a zero ion fraction
    x = gas_state(3, 4000, axes)
    assert np.array_equal(oracle.table_interp(table, axes, x), oracle.table_interp_scipy(table, axes, x), equal_nan=True)
