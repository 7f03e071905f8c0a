"""CPU check of the oracle against goldens produced by the REFERENCE'S OWN
kernels (compiled for sm_100 by oracle/build_ref.sh and run on a B200 by
tests/test_ref_parity.py::test_write_golden; the file is committed as
tests/golden/ref_golden.npz).  Input: img/bud_2 + img/bud_3, 640x384, D=64,
zd=32, ad_coeff=10, census_coeff=30, ucd=20, lcd=6, usd=17, lsd=9.

The golden also carries the GPU's two exponential tables (the only arithmetic
a CPU cannot reproduce bit for bit: ex2.approx); with them the oracle's whole
cost-volume chain must hash to the reference's outputs.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    path = os.path.join(GOLDEN, "ref_golden.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_golden.npz not generated yet")
    return np.load(path)


def test_oracle_matches_reference_goldens(oracle, golden, bud_sbs):
    H, W, D, ZD = [int(v) for v in golden["params"]]
    L = np.ascontiguousarray(bud_sbs[:, :W])
    R = np.ascontiguousarray(bud_sbs[:, W:])
    luts = (golden["lut_ad"], golden["lut_cen"])
    cl, cr = oracle.ci_adcensus(L, R, D, ZD, 10.0, 30.0, luts=luts)
    assert sha(cl) == str(golden["sha_cost_l"]) and sha(cr) == str(golden["sha_cost_r"])
    assert np.array_equal(cl[32, 100], golden["cost_l_d32_row100"])
    wta = {}
    for side, img, cost in (("l", L, cl), ("r", R, cr)):
        arms = oracle.cross_arms(img, 20.0, 6.0, 17, 9)
        assert sha(arms) == str(golden["sha_arms_" + side])
        acost = oracle.ca_aggregate(cost, arms)
        assert sha(acost) == str(golden["sha_acost_" + side])
        wta[side] = oracle.wta(acost, ZD)
        assert np.array_equal(wta[side].astype(np.int8), golden["wta_" + side])
        if side == "l":
            assert np.array_equal(acost[32, 100], golden["acost_l_d32_row100"])
    ol, orr = oracle.dcc(wta["l"], wta["r"])
    assert sha(ol) == str(golden["sha_outliers_l"]) and sha(orr) == str(golden["sha_outliers_r"])


def test_cpu_tables_within_tolerance_of_gpu_tables(oracle, golden):
    # the oracle's own exp2f tables vs the GPU's ex2.approx tables: <= 1e-5 on values in [0, 1]
    la, lc = oracle.exp_luts(10.0, 30.0)
    assert np.max(np.abs(la - golden["lut_ad"])) <= 2.5e-7
    assert np.max(np.abs(lc - golden["lut_cen"])) <= 2.5e-7
