"""GPU parity, stage by stage: every per-stage entry point of the C ABI
(include/s2mv.h, mirroring image_io.cpp:171-292) against the CPU oracle on
the same inputs.  Integer / byte / index stages and the exact-order fp32
aggregation are compared bit for bit; the only tolerance is the combine step
when the oracle uses its own CPU exponential tables (<= 1e-5 relative, the
bar BASELINE.md §5 states); with the GPU's tables it is bit-exact too.
"""
import numpy as np
import pytest

from conftest import DEFAULTS

pytestmark = pytest.mark.gpu

# (H, W, D, zd): in-domain reference shape, ragged shapes (W%160, W%32, H%32 != 0), D not a power of two,
# positive range larger than negative (no AD quirk), zero_disp = 1
SHAPES = [(96, 640, 64, 32), (50, 200, 20, 7), (37, 331, 12, 6), (33, 161, 24, 1), (40, 320, 128, 64)]


def pair(oracle, bud_sbs, H, W, seed=0):
    if W <= 640 and H <= 384:
        y0 = (seed * 37) % (384 - H + 1)
        x0 = (seed * 53) % (640 - W + 1)
        return (np.ascontiguousarray(bud_sbs[y0:y0 + H, x0:x0 + W]),
                np.ascontiguousarray(bud_sbs[y0:y0 + H, 640 + x0:640 + x0 + W]))
    from s2mv_b200_pkg import synth
    return synth.make_pair(H, W, 1000 + seed)


def test_gray_and_census_bit_exact(pipe, oracle, bud_sbs):
    for i, (H, W, _, _) in enumerate(SHAPES):
        L, _ = pair(oracle, bud_sbs, H, W, i)
        g = pipe.gray(L)
        assert np.array_equal(g, oracle.gray(L))
        assert np.array_equal(pipe.census(g), oracle.census(g))


@pytest.mark.parametrize("H,W,D,zd", SHAPES)
def test_ad_and_hamming_cost_bit_exact(pipe, oracle, bud_sbs, H, W, D, zd):
    L, R = pair(oracle, bud_sbs, H, W, 1)
    al, ar = pipe.ci_ad(L, R, D, zd)
    ol, orr = oracle.ad_cost(L, R, D, zd)
    assert np.array_equal(al, ol) and np.array_equal(ar, orr)
    hl, hr = pipe.ci_census(L, R, D, zd)
    CL, CR = oracle.census(oracle.gray(L)), oracle.census(oracle.gray(R))
    ol, orr = oracle.census_cost(CL, CR, D, zd)
    assert np.array_equal(hl, ol) and np.array_equal(hr, orr)


@pytest.mark.parametrize("H,W,D,zd", SHAPES)
def test_ci_adcensus(pipe, oracle, bud_sbs, H, W, D, zd):
    L, R = pair(oracle, bud_sbs, H, W, 2)
    cl, cr = pipe.ci_adcensus(L, R, 10.0, 30.0, D, zd)
    # (a) bit-exact when the oracle is fed the GPU's ex2.approx tables
    luts = pipe.exp_tables(10.0, 30.0)
    ol, orr = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0, luts=luts)
    assert np.array_equal(cl, ol) and np.array_equal(cr, orr)
    # (b) within 1e-5 relative of the CPU exp2f tables (values are in [0, 2])
    ol, orr = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0)
    assert np.max(np.abs(cl - ol)) <= 1e-5 * 2 and np.max(np.abs(cr - orr)) <= 1e-5 * 2


def test_exp_tables_close_to_cpu(pipe, oracle):
    for ad, cen in ((10.0, 30.0), (5.0, 12.5)):
        ga, gc = pipe.exp_tables(ad, cen)
        ca, cc = oracle.exp_luts(ad, cen)
        assert np.max(np.abs(ga - ca)) <= 2.5e-7 and np.max(np.abs(gc - cc)) <= 2.5e-7
        assert ga[0] == 0.0 and gc[0] == 0.0 and ga.min() >= 0.0 and ga.max() <= 1.0


def test_in_kernel_ad_term_equals_table(pipe):
    # the fused cost-initialisation kernel evaluates the AD exponential in place (ex2.approx.ftz, no
    # table); it must reproduce the reference-sequence table bit for bit for every possible sum
    for ad in (10.0, 5.0, 0.75, 30.0, 255.0, 1e-3):
        table, _ = pipe.exp_tables(ad, 30.0)
        assert np.array_equal(pipe.ad_terms(ad).view(np.uint32), table.view(np.uint32)), ad


@pytest.mark.parametrize("usd,lsd", [(0, 0), (1, 1), (5, 3), (33, 17), (40, 9)])
def test_ca_cross_halo_sizes(pipe, oracle, bud_sbs, usd, lsd):
    # arm caps other than the default 17/9: tile halo, segment planning and window bounds all depend on usd
    H, W, D, zd = 70, 333, 24, 12
    L, R = pair(oracle, bud_sbs, H, W, 5)
    cost, _ = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0)
    arms, acost = pipe.ca_cross(L, cost, 20.0, 6.0, usd, lsd)
    oarms = oracle.cross_arms(L, 20.0, 6.0, usd, lsd)
    assert np.array_equal(arms, oarms)
    assert np.array_equal(acost, oracle.ca_aggregate(cost, oarms))


@pytest.mark.parametrize("H,W,D,zd", SHAPES)
def test_ca_cross_bit_exact(pipe, oracle, bud_sbs, H, W, D, zd):
    L, R = pair(oracle, bud_sbs, H, W, 3)
    cost, _ = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0)
    arms, acost = pipe.ca_cross(L, cost, 20.0, 6.0, 17, 9)
    oarms = oracle.cross_arms(L, 20.0, 6.0, 17, 9)
    assert np.array_equal(arms, oarms)
    want = oracle.ca_aggregate(cost, oarms)
    assert np.array_equal(acost, want)          # sequential fp32 adds in the reference's order: exact


def test_ca_cross_other_arm_limits(pipe, oracle, bud_sbs):
    L, R = pair(oracle, bud_sbs, 60, 256, 4)
    cost, _ = oracle.ci_adcensus(L, R, 16, 8, 10.0, 30.0)
    # thresholds on both sides of 127 (packed-byte tests vs per-byte tests), fractional, negative, never failing
    for ucd, lcd, usd, lsd in ((30.0, 10.0, 25, 12), (5.0, 2.0, 4, 2), (20.0, 6.0, 1, 0), (127.0, 0.0, 17, 9),
                               (127.9, 126.5, 17, 9), (128.0, 6.0, 17, 9), (20.0, 200.0, 17, 9), (-1.0, 6.0, 17, 9),
                               (20.0, -0.5, 17, 9), (255.0, 300.0, 17, 9), (0.0, 0.0, 17, 9), (20.0, 6.0, 17, 40)):
        arms, acost = pipe.ca_cross(L, cost, ucd, lcd, usd, lsd)
        oarms = oracle.cross_arms(L, ucd, lcd, usd, lsd)
        assert np.array_equal(arms, oarms)
        assert np.array_equal(acost, oracle.ca_aggregate(cost, oarms))


def test_wta_first_minimum_bit_exact(pipe, oracle):
    r = np.random.default_rng(5)
    c = r.random((20, 33, 70), dtype=np.float32)
    c[7] = c[11]
    c[:, :5, :5] = 1.0                            # whole columns of ties -> d = 0
    assert np.array_equal(pipe.dc_wta(c, 9), oracle.wta(c, 9))


def stage_inputs(oracle, bud_sbs, H=96, W=320, D=32, zd=16):
    L, R = pair(oracle, bud_sbs, H, W, 6)
    dl, dr = oracle.costvol(L, R, D, zd)
    return L, R, dl, dr, D, zd


def test_dcc_irv_bit_exact(pipe, oracle, bud_sbs):
    L, R, dl, dr, D, zd = stage_inputs(oracle, bud_sbs)
    ol, orr = pipe.dr_dcc(dl, dr)
    wl, wr = oracle.dcc(dl, dr)
    assert np.array_equal(ol, wl) and np.array_equal(orr, wr)
    arms = oracle.cross_arms(L, 20.0, 6.0, 17, 9)
    for iters, host in ((1, True), (1, False), (5, False), (3, True)):
        gd, go = pipe.dr_irv(dl, wl, arms, 20, 0.4, D, zd, 17, iters, host_variant=host)
        od, oo = oracle.irv(dl, wl, arms, 20, 0.4, D, zd, 17, iters, host_variant=host)
        assert np.array_equal(gd, od) and np.array_equal(go, oo)
    # low thresholds so that most outliers get a vote
    gd, go = pipe.dr_irv(dl, wl, arms, 2, 0.01, D, zd, 17, 5)
    od, oo = oracle.irv(dl, wl, arms, 2, 0.01, D, zd, 17, 5)
    assert np.array_equal(gd, od) and np.array_equal(go, oo)
    assert (oo == 0).sum() > (wl == 0).sum()


def test_bilateral_bit_exact(pipe, oracle, bud_sbs):
    _, _, dl, _, D, _ = stage_inputs(oracle, bud_sbs)
    for radius, sc, ss in ((7, 5.0, 10.0), (7, 7.0, 7.0), (2, 3.0, 4.0)):
        assert np.array_equal(pipe.filter_bilateral_1(dl, radius, sc, ss, D), oracle.bilateral(dl, radius, sc, ss, D))


def test_occl_bleed_mask_bit_exact(pipe, oracle, bud_sbs):
    _, _, dl, dr, D, _ = stage_inputs(oracle, bud_sbs)
    fl, fr = oracle.bilateral(dl, 7, 5.0, 10.0, D), oracle.bilateral(dr, 7, 5.0, 10.0, D)
    ol, orr = pipe.dibr_occl(fl, fr)
    wl, wr = oracle.occl(fl, fr)
    assert np.array_equal(ol, wl) and np.array_equal(orr, wr)
    bl = pipe.filter_bleed_1(wl, 1)
    assert np.array_equal(bl, oracle.bleed(wl, 1))
    assert np.array_equal(pipe.filter_bleed_1(wr, 2), oracle.bleed(wr, 2))
    ml, mr = pipe.dibr_occl_to_mask(bl, wr)
    assert np.array_equal(ml, oracle.occl_to_mask(bl)) and np.array_equal(mr, oracle.occl_to_mask(wr))


def test_gaussian_dilate_bit_exact(pipe, oracle):
    r = np.random.default_rng(7)
    m = (r.random((70, 90)) < 0.9).astype(np.float32)
    for radius, sigma in ((10, 15.0), (7, 10.0)):
        assert np.array_equal(pipe.filter_gaussian_1(m, radius, sigma), oracle.gaussian_dilate(m, radius, sigma))
    f = r.random((45, 61), dtype=np.float32)
    assert np.array_equal(pipe.filter_gaussian_1(f, 3, 2.0), oracle.gaussian_dilate(f, 3, 2.0))


def test_dbm_and_mux_bit_exact(pipe, oracle, bud_sbs):
    L, R, dl, dr, D, _ = stage_inputs(oracle, bud_sbs)
    fl, fr = oracle.bilateral(dl, 7, 5.0, 10.0, D), oracle.bilateral(dr, 7, 5.0, 10.0, D)
    ol, orr = oracle.occl(fl, fr)
    ml, mr = oracle.occl_to_mask(oracle.bleed(ol, 1)), oracle.occl_to_mask(oracle.bleed(orr, 1))
    views = [R]
    for v in range(1, 7):
        shift = np.float32(1.0 - (1.0 * v) / 7.0)
        got = pipe.dibr_dbm(L, R, fl, fr, ml, mr, shift, 10, 15.0)
        assert np.array_equal(got, oracle.dbm(L, R, fl, fr, ml, mr, shift, 10, 15.0)), v   # warped indices exact
        views.append(got)
    views.append(L)
    got7 = pipe.dibr_dbm(L, R, fl, fr, ml, mr, 0.5, 7, 10.0)            # the image path's blur (d_dibr_bwarp.cu:151)
    assert np.array_equal(got7, oracle.dbm(L, R, fl, fr, ml, mr, 0.5, 7, 10.0))
    H, W, _ = L.shape
    for (Ho, Wo, variant, angle) in ((H, W, 2, 18.0), (H, W, 1, 18.0), (120, 400, 0, 18.0), (101, 333, 0, 23.0)):
        got = pipe.mux_multiview(views, angle, Ho, Wo, variant)
        kv = variant if variant else (2 if Ho % 8 == 0 else 1)
        assert np.array_equal(got, oracle.mux_multiview(views, angle, Ho, Wo, kv)), (Ho, Wo, variant)


def test_bad_parameters_are_rejected(s2mv, pipe):
    L = np.zeros((8, 8, 3), np.uint8)
    with pytest.raises(s2mv.S2mvError):
        pipe.configure(num_rows=8, num_cols=8, elem_sz=4)
    with pytest.raises(s2mv.S2mvError):
        pipe.configure(num_rows=8, num_cols=8, num_views=1)
    with pytest.raises(s2mv.S2mvError):
        pipe.configure(num_rows=8, num_cols=8, angle=0)      # zero interlace period: the reference divides by zero (Q27)
    with pytest.raises(s2mv.S2mvError):
        pipe.filter_bilateral_1(np.zeros((8, 8), np.float32), 99, 1.0, 1.0, 8)
    p2 = s2mv.Pipeline(0)
    with pytest.raises(s2mv.S2mvError, match="configure"):
        p2.process_device(1, 16)
    p2.close()
    assert pipe.gray(L).shape == (8, 8)
