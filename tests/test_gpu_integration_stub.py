"""-m gpu: the reference-side binding shown in INTEGRATION.md section 3 is executed VERBATIM (the two python blocks, with the
library path filled in) and checked against the oracle / scipy -- the document cannot drift from the C ABI."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_l2, random_cloud

pytestmark = pytest.mark.gpu


def _stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 3."):text.index("## 4.")]
    blocks = re.findall(r"```python\n(.*?)```", sec, flags=re.S)
    assert len(blocks) == 2
    from astro_sph_tools_b200 import _lib
    from astro_sph_tools_b200.tools.projections import quartic_spline_kernel
    src = "\n".join(blocks).replace('"libastsph_b200.so"', repr(_lib.LIB_PATH))
    ns = {"quartic_spline_kernel": quartic_spline_kernel}
    exec(compile(src, "INTEGRATION.md#3", "exec"), ns)
    return ns


def test_documented_create_image_stub_runs_and_matches_the_oracle(oracle):
    from astro_sph_tools_b200 import CoordinateAxes
    ns = _stub_namespace()
    pos, h, prop = random_cloud(4, 4000, h_hi=0.6)
    img = ns["create_image"](pos, h, prop, (160, 160), 32, CoordinateAxes.Z, 0.0, 10.0, 0.0, 10.0)
    ref = oracle.project2d(pos, h, prop, (160, 160), 2, 0.0, 10.0, 0.0, 10.0)
    assert img.shape == (160, 160) and img.dtype == np.float64
    assert rel_l2(img, ref) <= 1e-5 and abs(img.sum() - ref.sum()) <= 1e-6 * np.abs(ref).sum()


def test_documented_knn_stub_is_bit_equal_to_scipy():
    from scipy.spatial import cKDTree
    ns = _stub_namespace()
    pos = np.random.default_rng(3).uniform(0, 1, (20000, 3))
    h = ns["knn_smoothing_lengths"](pos, 32)
    assert np.array_equal(h, cKDTree(pos).query(pos, k=32)[0][:, 31])
