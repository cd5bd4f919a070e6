"""CPU: the k-NN restatement in the oracle (orc_knn_brute) is pinned bit-for-bit against scipy.spatial.cKDTree,
which is the arithmetic the reference actually calls (io/SWIFT/_SnapshotSWIFT.py:69-82; scipy is third-party,
un-pinned by the reference, 1.18.1 in this image)."""
import numpy as np
import pytest


@pytest.mark.parametrize("box", [None, 1.0])
def test_brute_force_restatement_bit_equals_scipy(oracle, box):
    rng = np.random.default_rng(4)
    pos = rng.uniform(0, 1.0, (1500, 3))
    h_s, d_s, i_s = oracle.knn_scipy(pos, 32, box)
    h_o, d_o, i_o = oracle.knn_brute(pos, 32, box or 0.0, want_lists=True)
    assert np.array_equal(h_o, h_s)                   # bitwise
    assert np.array_equal(d_o, d_s)
    assert np.array_equal(i_o, i_s)                   # distinct distances -> identical neighbour lists
    assert np.all(d_s[:, 0] == 0.0) and np.array_equal(i_s[:, 0], np.arange(1500))   # self is neighbour #1


def test_fewer_points_than_k_gives_inf(oracle):
    pos = np.random.default_rng(1).uniform(0, 1, (10, 3))
    assert np.all(np.isinf(oracle.knn_brute(pos, 32)))
    assert np.all(np.isinf(oracle.knn_scipy(pos, 32)[0]))
