import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the reference documents no defaults; these are BASELINE.md §4's
DEFAULTS = dict(ad_coeff=10.0, census_coeff=30.0, ucd=20.0, lcd=6.0, usd=17, lsd=9, thresh_s=20, thresh_h=0.4,
                num_views=8, angle=18)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_sbs(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))["sbs"]


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def bud_sbs():
    return load_sbs("bud_2_3")


@pytest.fixture(scope="session")
def fish_sbs():
    return load_sbs("fish_1_2")


@pytest.fixture(scope="session")
def s2mv():
    import s2mv_b200
    s2mv_b200.build()
    return s2mv_b200


@pytest.fixture(scope="session")
def pipe(s2mv):
    """A context on cuda:0 (GPU tests only)."""
    p = s2mv.Pipeline(0)
    yield p
    p.close()
