"""Scanline optimisation (the stage the reference declares as dc_hslo and never finished).
PARITY UNPINNED against the reference — there is no reference output; oracle/s2mv_oracle.c:orc_so is the
specification (Mei et al. 2011 + the stub's constants / colour measure / tiers) and the CUDA kernels are held
to it bit for bit.  CPU: properties of the specification.  GPU: kernels vs specification."""
import numpy as np
import pytest

from conftest import DEFAULTS

ALGO = {k: DEFAULTS[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}


def _aggregated(oracle, H, W, D, zd, seed):
    from s2mv_b200_pkg import synth
    sbs = synth.make_sbs(H, W, seed)
    L, R = np.ascontiguousarray(sbs[:, :W]), np.ascontiguousarray(sbs[:, W:])
    cl, cr = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0)
    al = oracle.ca_aggregate(cl, oracle.cross_arms(L, 20.0, 6.0, 17, 9))
    ar = oracle.ca_aggregate(cr, oracle.cross_arms(R, 20.0, 6.0, 17, 9))
    return L, R, al, ar


def test_specification_properties(s2mv, oracle):
    L, R, al, ar = _aggregated(oracle, 36, 72, 24, 10, 11)
    # zero penalties: every direction reproduces the cost up to the rounding of (C + m) - m, so the result is
    # plain winner-takes-all (but for exact ties broken by that rounding)
    d0, c0 = oracle.so(al, L, R, 0, 15.0, 0.0, 0.0, 10, want_cost=True)
    assert np.allclose(c0, al, rtol=1e-4, atol=1e-2) and (d0 == oracle.wta(al, 10)).mean() > 0.999
    # the optimised cost never drops below the cost and exceeds it by at most the largest jump penalty
    d1, c1 = oracle.so(al, L, R, 0, 15.0, 1.0, 3.0, 10, want_cost=True)
    assert (c1 >= al - 1e-2).all() and (c1 <= al + 3.0 + 1e-2).all()      # up to the rounding of (C + best) - m
    assert d1.min() >= -10 and d1.max() <= 24 - 1 - 10
    # a volume that prefers one disparity everywhere keeps it
    flat = np.ones_like(al)
    flat[7] = 0.5
    assert (oracle.so(flat, L, R, 0, 15.0, 1.0, 3.0, 10) == 7 - 10).all()
    # huge penalties: each scanline is forced to (nearly) one disparity, so far fewer disparity changes along rows
    d2 = oracle.so(al, L, R, 0, 15.0, 1e6, 3e6, 10)
    assert (np.diff(d2, axis=1) != 0).mean() < (np.diff(oracle.wta(al, 10), axis=1) != 0).mean()
    # right view runs with the mirrored shift
    dr = oracle.so(ar, R, L, 1, 15.0, 1.0, 3.0, 10)
    assert dr.shape == d1.shape


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,D,zd", [(40, 72, 24, 10), (33, 50, 37, 0), (24, 160, 128, 64), (17, 35, 5, 4)])
def test_kernels_match_specification(s2mv, oracle, pipe, H, W, D, zd):
    L, R, al, ar = _aggregated(oracle, H, W, D, zd, 100 + D)
    for view, (cost, own, oth) in enumerate(((al, L, R), (ar, R, L))):
        want_d, want_c = oracle.so(cost, own, oth, view, 15.0, 1.0, 3.0, zd, want_cost=True)
        got_d, got_c = pipe.dc_so(cost, own, oth, view, 15.0, 1.0, 3.0, zd, want_cost=True)
        assert np.array_equal(got_c, want_c), f"view {view}: optimised cost"
        assert np.array_equal(got_d, want_d), f"view {view}: disparities"
    # other penalties / threshold, disparities only
    assert np.array_equal(pipe.dc_so(al, L, R, 0, 4.0, 50.0, 700.0, zd), oracle.so(al, L, R, 0, 4.0, 50.0, 700.0, zd))
    # zero penalties = plain winner-takes-all (up to ties broken by the rounding of (C + m) - m)
    assert (pipe.dc_so(al, L, R, 0, 15.0, 0.0, 0.0, zd) == pipe.dc_wta(al, zd)).mean() > 0.999


@pytest.mark.gpu
def test_frame_path_with_scanline_optimisation(s2mv, oracle):
    from s2mv_b200_pkg import synth
    H, W, D, zd = 64, 160, 32, 16
    sbs = synth.make_sbs(H, W, 2024)
    L, R = np.ascontiguousarray(sbs[:, :W]), np.ascontiguousarray(sbs[:, W:])
    with s2mv.Pipeline(0, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO) as p:
        plain = p.adcensus_stm(sbs)
        luts = p.exp_tables()
        p.enable_so(True, 15.0, 1.0, 3.0)
        p.enable_taps(True)
        dl, dr, out = p.adcensus_stm(sbs)
        taps = p.read_taps()
        again = p.adcensus_stm(sbs)
        p.enable_so(False)
        back = p.adcensus_stm(sbs)
    o = oracle.adcensus_stm(sbs, W, H, W, num_views=8, angle=18, D=D, zd=zd, luts=luts, want_taps=True, **ALGO)[3]
    assert np.array_equal(taps["wta_l"], oracle.so(o["acost_l"], L, R, 0, 15.0, 1.0, 3.0, zd))
    assert np.array_equal(taps["wta_r"], oracle.so(o["acost_r"], R, L, 1, 15.0, 1.0, 3.0, zd))
    assert all(np.array_equal(a, b) for a, b in zip((dl, dr, out), again))      # deterministic
    assert all(np.array_equal(a, b) for a, b in zip(plain, back))               # switching it off restores adcensus_stm
    with s2mv.Pipeline(0, num_rows=H, num_cols=W, num_disp=200, zero_disp=100, **ALGO) as p:
        with pytest.raises(s2mv.S2mvError):
            p.enable_so(True)                                                    # num_disp > 128
