import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")
    config.addinivalue_line("markers", "slow: longer CPU test")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not f.endswith("_summary.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.lib()          # builds oracle/liboracle.so with gcc if needed
    return orc


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


def random_cloud(seed, n, L=10.0, h_lo=0.0, h_hi=1.0, signed=False):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, L, (n, 3))
    h = rng.uniform(h_lo, h_hi, n)
    prop = rng.normal(size=n) if signed else rng.uniform(0.5, 1.5, n)
    return pos, h, prop
