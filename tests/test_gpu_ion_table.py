"""-m gpu: table interpolation (ast_table_interp through the IonisationTableBase mirror) against scipy's
RegularGridInterpolator -- the arithmetic the reference calls (data_structures/_IonisationTable.py:44-58) -- bit for bit,
against golden values from the reference's own class, and the fused ion column-density path against the two-step one."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, rel_l2
from table_util import gas_state, synthetic_table

pytestmark = pytest.mark.gpu


def _table(ndim, uniform, shape=(41, 141, 49)):
    table, axes = synthetic_table(uniform=uniform, shape=shape)
    sl = (slice(None),) * ndim + (3,) * (3 - ndim)
    return np.ascontiguousarray(table[sl]), axes[:ndim]


@pytest.mark.parametrize("uniform", [True, False])
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_call_bit_equals_scipy(oracle, ndim, uniform):
    from astro_sph_tools_b200.tools.ionisation import IonisationTableBase
    table, axes = _table(ndim, uniform)
    t = IonisationTableBase(table, *axes, redshift_input_index=ndim - 1)
    x = gas_state(100 + ndim, 200_000, axes)
    got = t(x)
    ref = oracle.table_interp_scipy(table, axes, x)
    assert got.shape == ref.shape and got.dtype == np.float64
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.isnan(got[40]) and np.isneginf(got[44])


def test_evaluate_at_redshift_bit_equals_scipy_and_reference_golden(oracle):
    from astro_sph_tools_b200.tools.ionisation import IonisationTableBase
    g = np.load(os.path.join(GOLDEN_DIR, "tables", "ion_table_golden.npz"))
    axes = [g["axis0"], g["axis1"], g["axis2"]]
    t = IonisationTableBase(g["table"], *axes, redshift_input_index=2)
    assert np.array_equal(t(g["x"]), g["call"], equal_nan=True)                                        # reference class output
    assert np.array_equal(t.evaluate_at_redshift(g["x2"], float(g["redshift"])), g["at_redshift"], equal_nan=True)
    table, axes = _table(3, True)
    t = IonisationTableBase(table, *axes, redshift_input_index=2)
    x2 = np.ascontiguousarray(gas_state(5, 100_000, axes)[:, :2])
    for z in (0.0, 0.1, 3.3, axes[2][-1], 9.5, -0.1):
        ref = oracle.table_interp_scipy(table, axes, np.insert(x2, 2, z, axis=1))
        assert np.array_equal(t.evaluate_at_redshift(x2, z), ref, equal_nan=True), z
    assert t.number_of_input_dimensions == 3 and np.array_equal(t.ionisation_fraction_table, table)
    assert np.array_equal(t.get_table_dimension(1), axes[1])


def test_infinite_table_entries_and_4d(oracle):
    from astro_sph_tools_b200.tools.ionisation import IonisationTableBase
    table, axes = synthetic_table(shape=(9, 11, 5))
    table[2, 3, 1] = -np.inf
    x = gas_state(3, 50_000, axes)
    assert np.array_equal(IonisationTableBase(table, *axes)(x), oracle.table_interp_scipy(table, axes, x), equal_nan=True)
    rng = np.random.default_rng(0)
    axes4 = [np.sort(rng.uniform(0, 1, n)) for n in (5, 6, 7, 4)]
    t4 = rng.normal(size=(5, 6, 7, 4))
    x4 = rng.uniform(-0.05, 1.05, (30_000, 4))
    assert np.array_equal(IonisationTableBase(t4, *axes4)(x4), oracle.table_interp_scipy(t4, axes4, x4), equal_nan=True)
    # a large axis (> 4096 grid points in total) takes the global-memory axis path
    big_axis = [np.linspace(0, 1, 6000)]
    tb = rng.normal(size=6000)
    xb = rng.uniform(-0.01, 1.01, (40_000, 1))
    assert np.array_equal(IonisationTableBase(tb, *big_axis)(xb), oracle.table_interp_scipy(tb, big_axis, xb), equal_nan=True)


def test_errors_mirror_reference():
    from astro_sph_tools_b200.tools.ionisation import IonisationTableBase
    with pytest.raises(IndexError, match="No input dimensions"):
        IonisationTableBase(np.zeros((3, 3)))
    with pytest.raises(IndexError, match="Interpolation table has 2 dimensions but 1 arrays"):
        IonisationTableBase(np.zeros((3, 3)), np.arange(3.0))
    t = IonisationTableBase(np.zeros((3, 3)), np.arange(3.0), np.arange(3.0))
    with pytest.raises(ValueError, match="dimension 3"):
        t(np.zeros((4, 3)))


def test_fused_ion_weights_and_column_map(oracle):
    """weights = mass * 10**table(lognH, logT, z) (bit-equal interpolation, exp10 within 4 ulp); the fused map equals the
    two-step create_image(weights from scipy) within the projection gates"""
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.ionisation import IonisationTableBase, create_ion_image, ion_weights
    from astro_sph_tools_b200.tools.projections import create_image
    table, axes = _table(3, True)
    t = IonisationTableBase(table, *axes, redshift_input_index=2)
    rng = np.random.default_rng(9)
    n = 60_000
    pos = rng.uniform(0, 1, (n, 3)); h = rng.uniform(0.002, 0.03, n); m = rng.uniform(0.5, 1.5, n)
    lognh = rng.uniform(-8.3, 2.2, n); logt = rng.uniform(1.9, 9.1, n); z = 2.2
    ref_w = m * 10.0 ** oracle.table_interp_scipy(table, axes, np.stack([lognh, logt, np.full(n, z)], axis=1))
    w = ion_weights(t, m, lognh, logt, z)
    assert np.all(w[ref_w == 0] == 0) and (ref_w == 0).sum() > 100            # outside the table: weight exactly 0
    np.testing.assert_allclose(w, ref_w, rtol=1e-15 * 8, atol=0)
    args = ((256, 256), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0)
    fused = create_ion_image(t, pos, h, m, lognh, logt, z, *args)
    two_step = create_image(pos, h, ref_w, *args)
    ref = oracle.project2d(pos, h, ref_w, (256, 256), 2, 0.0, 1.0, 0.0, 1.0)
    assert rel_l2(fused, two_step) <= 1e-12
    assert rel_l2(fused, ref) <= 1e-5 and abs(fused.sum() - ref.sum()) <= 1e-6 * abs(ref.sum())
