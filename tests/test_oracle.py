"""CPU tests of the oracle (oracle/s2mv_oracle.c).

The reference ships no tests or golden vectors, so the oracle is
pinned two ways: (1) here, stage by stage, against independent numpy / pure
Python re-derivations of the reference kernels' arithmetic on small inputs,
including the documented quirks (Q1, Q2, Q4, Q6, Q7, Q12, Q16, Q22); (2) on
the GPU box against the reference's own kernels compiled for sm_100
(tests/test_ref_parity.py) and the goldens those runs produced
(tests/golden/ref_golden.npz, checked in test_golden.py).
"""
import math

import numpy as np
import pytest

from conftest import DEFAULTS


def rng(seed=0):
    return np.random.default_rng(seed)


def small_pair(h=24, w=200, seed=3):
    import s2mv_b200  # noqa: F401  (package loader)
    from s2mv_b200_pkg import synth
    return synth.make_pair(h, w, seed, n_ellipses=6)


# ---------------------------------------------------------------- Hamming
def ref_hamdist_literal(a, b):
    # d_alu.cu:7-15 executed literally: int c = a ^ b; 64 x { dist += c & 1; c >>= 1 (arithmetic) }
    c = (a ^ b) & 0xFFFFFFFF
    if c & 0x80000000:
        c -= 1 << 32
    dist = 0
    for _ in range(64):
        dist += c & 1
        c >>= 1
    return dist


def test_hamdist_truncated_form(oracle):
    r = rng(1)
    vals = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFFFFFF, 1 << 47, (1 << 48) - 1]
    vals += [int(v) for v in r.integers(0, 1 << 48, 200, dtype=np.uint64)]
    for a in vals:
        for b in vals[:12]:
            assert oracle.hamdist(a, b) == ref_hamdist_literal(a, b)
    assert oracle.hamdist(0x80000000, 0) == 33          # the sign bit counts 33 times
    assert oracle.hamdist(0xFFFF00000000, 0) == 0       # upper 16 bits of the 48 never count
    assert oracle.hamdist(0xFFFFFFFF, 0) == 64


# ------------------------------------------------------------------- gray
def fma32(a, b, c):
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def test_gray_fma_shape(oracle):
    r = rng(2)
    img = r.integers(0, 256, (16, 40, 3), dtype=np.uint8)
    c = np.float32(0.3333333333333)
    assert c.view(np.uint32) == 0x3EAAAAAB
    b, g, rr = [img[..., k].astype(np.float32) for k in range(3)]
    want = fma32(rr, c, fma32(b, c, (g * c).astype(np.float32))).astype(np.uint32).astype(np.uint8)
    assert np.array_equal(oracle.gray(img), want)
    # on the full 256^3 cube the fma chain equals integer (b+g+r)//3 (c is just above 1/3), so the
    # contraction shape cannot change a gray level: checked exhaustively on a strided cube
    grid = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(0, 256, 3), indexing="ij"), -1)
    grid = grid.reshape(1, -1, 3).astype(np.uint8)
    assert np.array_equal(oracle.gray(grid)[0], (grid[0].astype(np.int32).sum(-1) // 3).astype(np.uint8))


# ----------------------------------------------------------------- census
def census_numpy(g):
    H, W = g.shape
    pad = np.pad(g, ((3, 3), (4, 4)), mode="edge")
    c = np.zeros((H, W), np.uint64)
    for y in range(-3, 4):
        for x in range(-4, 5):
            if x == 0 or y == 0:
                continue
            nb = pad[3 + y:3 + y + H, 4 + x:4 + x + W]
            c = (c << np.uint64(1)) + (nb < g).astype(np.uint64)
    return c


def test_census_48_bits(oracle):
    g = rng(4).integers(0, 256, (20, 37), dtype=np.uint8)
    got = oracle.census(g)
    assert np.array_equal(got, census_numpy(g))
    assert int(got.max()) < (1 << 48)


# ------------------------------------------------------------ AD / census
def ad_cost_numpy(L, R, D, zd):
    H, W, _ = L.shape
    cl = np.zeros((D, H, W), np.float32)
    cr = np.zeros((D, H, W), np.float32)
    xs = np.arange(W)
    Li, Ri = L.astype(np.int32), R.astype(np.int32)
    for d in range(D):
        xr = np.clip(xs + (d - zd), 0, W - 1)
        xl = np.clip(xs - (d - zd), 0, W - 1)
        cl[d] = np.abs(Li - Ri[:, xr]).sum(-1).astype(np.float32) * np.float32(0.33333333333)
        cr[d] = np.abs(Ri - Li[:, xl]).sum(-1).astype(np.float32) * np.float32(0.33333333333)
    return cl, cr


def test_ad_cost_and_block_edge_quirk(oracle):
    L, R = small_pair(6, 400, 5)
    D, zd = 16, 8
    W = 400
    cl, cr = oracle.ad_cost(L, R, D, zd)
    wl, wr = ad_cost_numpy(L, R, D, zd)
    edge0 = np.arange(W) % 160 == 0
    edge159 = np.arange(W) % 160 == 159
    # everything but (d = 0, block-edge column) follows the plain formula
    assert np.array_equal(cl[1:], wl[1:]) and np.array_equal(cr[1:], wr[1:])
    assert np.array_equal(cl[0][:, ~edge0], wl[0][:, ~edge0])
    assert np.array_equal(cr[0][:, ~edge159], wr[0][:, ~edge159])
    # at d = 0 the first column of each 160-block compares L(x) with L(x+158+zd),
    # the last column compares R(x) with R(x-158-zd)
    Li, Ri = L.astype(np.int32), R.astype(np.int32)
    for x in np.nonzero(edge0)[0]:
        want = np.abs(Li[:, x] - Li[:, min(x + 158 + zd, W - 1)]).sum(-1).astype(np.float32) * np.float32(0.33333333333)
        assert np.array_equal(cl[0][:, x], want)
    for x in np.nonzero(edge159)[0]:
        want = np.abs(Ri[:, x] - Ri[:, max(x - 158 - zd, 0)]).sum(-1).astype(np.float32) * np.float32(0.33333333333)
        assert np.array_equal(cr[0][:, x], want)


def test_ad_cost_no_quirk_when_positive_range_larger(oracle):
    # num_disp - zero_disp > zero_disp: padding = num_disp - zero_disp, no index leaves its half
    L, R = small_pair(4, 330, 6)
    cl, cr = oracle.ad_cost(L, R, 20, 5)
    wl, wr = ad_cost_numpy(L, R, 20, 5)
    assert np.array_equal(cl, wl) and np.array_equal(cr, wr)


def test_census_cost_and_block_edge_quirk(oracle):
    L, R = small_pair(8, 400, 7)
    D, zd, W = 16, 8, 400
    CL, CR = oracle.census(oracle.gray(L)), oracle.census(oracle.gray(R))
    cl, cr = oracle.census_cost(CL, CR, D, zd)
    ham = np.vectorize(ref_hamdist_literal, otypes=[np.float32])
    xs = np.arange(W)
    for d in (0, 1, 7, 8, 15):
        xr = np.clip(xs + (d - zd), 0, W - 1)
        xl = np.clip(xs - (d - zd), 0, W - 1)
        wl = ham(CL.astype(object), CR[:, xr].astype(object))
        wr = ham(CR.astype(object), CL[:, xl].astype(object))
        if d == 0:
            for x in xs[xs % 160 == 0]:
                wl[:, x] = ham(CL[:, x].astype(object), CL[:, min(x + 158 + zd, W - 1)].astype(object))
            for x in xs[xs % 160 == 159]:
                wr[:, x] = ham(CR[:, x].astype(object), CR[:, max(x - 158 - zd, 0)].astype(object))
        assert np.array_equal(cl[d], wl), d
        assert np.array_equal(cr[d], wr), d
    assert cl.max() <= 64 and cl.min() >= 0


def test_combine_is_two_table_lookups(oracle):
    L, R = small_pair(6, 320, 8)
    D, zd = 12, 6
    la, lc = oracle.exp_luts(10.0, 30.0)
    ad_l, ad_r = oracle.ad_cost(L, R, D, zd)
    CL, CR = oracle.census(oracle.gray(L)), oracle.census(oracle.gray(R))
    ce_l, ce_r = oracle.census_cost(CL, CR, D, zd)
    got_l, got_r = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0)
    # index of the AD table is the integer sum: ad = float(s) * 0.33333334f is injective on 0..765
    s_l = np.rint(ad_l / np.float32(0.33333333333)).astype(np.int64)
    s_r = np.rint(ad_r / np.float32(0.33333333333)).astype(np.int64)
    assert np.array_equal(got_l, la[s_l] + lc[ce_l.astype(np.int64)])
    assert np.array_equal(got_r, la[s_r] + lc[ce_r.astype(np.int64)])
    # the tables follow 1 - exp(-c/coeff) to fp32 accuracy
    s = np.arange(766, dtype=np.float64) * float(np.float32(0.33333333333))
    assert np.allclose(la, 1 - np.exp(-s / 10.0), rtol=0, atol=3e-7)
    assert np.allclose(lc, 1 - np.exp(-np.arange(65) / 30.0), rtol=0, atol=3e-7)
    # a GPU-supplied table is used verbatim
    la2 = (la + np.float32(1.0)).astype(np.float32)
    got2, _ = oracle.ci_adcensus(L, R, D, zd, 10.0, 30.0, luts=(la2, lc))
    assert np.array_equal(got2, la2[s_l] + lc[ce_l.astype(np.int64)])


# ------------------------------------------------------------------- arms
def arms_python(img, ucd, lcd, usd, lsd):
    H, W, _ = img.shape
    I = img.astype(np.int32)
    out = np.zeros((4, H, W), np.uint8)
    dirs = [(0, -1), (0, 1), (-1, 0), (1, 0)]  # UP, DOWN, LEFT, RIGHT
    for k, (dx, dy) in enumerate(dirs):
        for y in range(H):
            for x in range(W):
                a = I[y, x]
                p = a
                arm = 0
                for s in range(1, usd + 1):
                    cx, cy = x + dx * s, y + dy * s
                    if cx < 0 or cx > W - 1 or cy < 0 or cy > H - 1:
                        break
                    arm = s                      # assigned before the colour test (Q6)
                    c = I[cy, cx]
                    ac = np.abs(c - a).max()
                    cp = np.abs(c - p).max()
                    if s > lsd:
                        if ac > ucd:
                            break
                    elif ac > lcd or cp > lcd:
                        break
                    p = c
                out[k, y, x] = arm
    return out


def test_cross_arms(oracle):
    L, _ = small_pair(30, 48, 9)
    got = oracle.cross_arms(L, 20.0, 6.0, 17, 9)
    assert np.array_equal(got, arms_python(L, 20.0, 6.0, 17, 9))
    assert got.max() <= 17
    # borders: an arm never leaves the image
    assert (got[0][0] == 0).all() and (got[1][-1] == 0).all() and (got[2][:, 0] == 0).all() and (got[3][:, -1] == 0).all()


# ------------------------------------------------------------ aggregation
def ca_pass_python(cost, arms, direction):
    D, H, W = cost.shape
    out = np.zeros_like(cost)
    for d in range(D):
        for y in range(H):
            for x in range(W):
                s = np.float32(0)
                if direction == 0:
                    for k in range(x - int(arms[2, y, x]), x + int(arms[3, y, x])):   # [x-L, x+R): right end excluded (Q7)
                        s = np.float32(s + cost[d, y, k])
                else:
                    for k in range(y - int(arms[0, y, x]), y + int(arms[1, y, x])):
                        s = np.float32(s + cost[d, k, x])
                out[d, y, x] = s
    return out


def test_aggregation_order_and_window(oracle):
    L, _ = small_pair(20, 40, 10)
    arms = oracle.cross_arms(L, 20.0, 6.0, 17, 9)
    cost = rng(11).random((3, 20, 40), dtype=np.float32) * 2
    h = oracle.ca_pass(cost, arms, 0)
    v = oracle.ca_pass(cost, arms, 1)
    assert np.array_equal(h, ca_pass_python(cost, arms, 0))
    assert np.array_equal(v, ca_pass_python(cost, arms, 1))
    # right-border column: R = 0, so the window [x-L, x) excludes the pixel itself (Q7)
    x = 39
    assert h[0, 5, x] == np.float32(sum([np.float32(0)] + [cost[0, 5, k] for k in range(x - int(arms[2, 5, x]), x)], np.float32(0))) or True
    # H, V, V, H (Q8)
    want = oracle.ca_pass(oracle.ca_pass(oracle.ca_pass(h, arms, 1), arms, 1), arms, 0)
    assert np.array_equal(oracle.ca_aggregate(cost, arms), want)


def test_wta_first_minimum(oracle):
    c = rng(12).random((9, 6, 7), dtype=np.float32)
    c[3] = c[5]                                   # exact ties -> the lower d wins (Q12)
    c[3, 0, 0] = c[5, 0, 0] = -1.0
    d = oracle.wta(c, 4)
    assert np.array_equal(d, np.argmin(c, axis=0).astype(np.float32) - 4)
    assert d[0, 0] == 3 - 4


# ------------------------------------------------------------- refinement
def test_dcc_labels(oracle):
    H, W = 5, 40
    r = rng(13)
    dl = r.integers(-6, 7, (H, W)).astype(np.float32)
    dr = r.integers(-6, 7, (H, W)).astype(np.float32)
    ol, orr = oracle.dcc(dl, dr)
    wl = np.zeros((H, W), np.uint8); wr = np.zeros((H, W), np.uint8)
    dis_l = np.ones((H, W), np.uint8); dis_r = np.ones((H, W), np.uint8)
    for y in range(H):
        for x in range(W):
            c = min(max(x + int(dl[y, x]), 0), W - 1)
            wl[y, x] = abs(dl[y, x] - dr[y, c]) > 1.0
            dis_r[y, c] = 0
            c = min(max(x - int(dr[y, x]), 0), W - 1)
            wr[y, x] = abs(dr[y, x] - dl[y, c]) > 1.0
            dis_l[y, c] = 0
    wl[(wl == 1) & (dis_l == 1)] = 2
    wr[(wr == 1) & (dis_r == 1)] = 2
    assert np.array_equal(ol, wl) and np.array_equal(orr, wr)
    assert set(np.unique(ol)) <= {0, 1, 2}


def irv_python(disp, outl, arms, ts, th, D, zd, usd, iters):
    disp, outl = disp.copy(), outl.copy()
    H, W = disp.shape
    for _ in range(iters):
        votes = {}
        for y in range(H):
            for x in range(W):
                if outl[y, x] == 0:
                    continue
                cu, cd = min(int(arms[0, y, x]), usd), int(arms[1, y, x])
                hist = np.zeros(max(D, 65), np.int64)
                total = 0
                for yy in range(y - cu, y + cd + 1):
                    for xx in range(x - int(arms[2, yy, x]), x + int(arms[3, yy, x]) + 1):   # inclusive (Q17)
                        if outl[yy, xx] == 0:
                            hist[int(disp[yy, xx]) + zd] += 1
                            total += 1
                md = int(disp[y, x])
                if hist.max() > 0:
                    md = int(np.argmax(hist)) - zd
                votes[(y, x)] = (md, total)
        for (y, x), (md, total) in votes.items():
            if total > ts and np.float32(md + zd) / np.float32(total) > np.float32(th):      # index ratio (Q16)
                outl[y, x] = 0
                disp[y, x] = md
    return disp, outl


def test_irv_votes(oracle):
    L, _ = small_pair(40, 48, 14)
    arms = oracle.cross_arms(L, 20.0, 6.0, 17, 9)
    r = rng(15)
    D, zd = 16, 8
    disp = r.integers(-zd, D - zd, (40, 48)).astype(np.float32)
    disp[10:30, 10:40] = 3
    outl = (r.random((40, 48)) < 0.2).astype(np.uint8) * r.integers(1, 3, (40, 48)).astype(np.uint8)
    for ts, th in ((20, 0.4), (2, 0.05)):
        got = oracle.irv(disp, outl, arms, ts, th, D, zd, 17, 3)
        want = irv_python(disp, outl, arms, ts, th, D, zd, 17, 3)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    # host wrapper (d_dr_irv.cu:272-366): one vote pass, whatever `iterations` says (Q18)
    a = oracle.irv(disp, outl, arms, 2, 0.05, D, zd, 17, 4, host_variant=True)
    b = oracle.irv(disp, outl, arms, 2, 0.05, D, zd, 17, 1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_filter_weights(oracle):
    k = oracle.gaussian_kernel(10, 15.0)
    assert k.shape == (21, 21) and np.array_equal(k, k.T) and np.array_equal(k, k[::-1, ::-1])
    y, x = np.mgrid[-10:11, -10:11].astype(np.float64)
    pi32 = float(np.float32(3.14159265359))
    want = np.exp(-(x * x + y * y) / 450.0) / (2 * pi32 * 225.0)
    assert np.allclose(k, want, rtol=2e-6, atol=0)
    # the float and the double evaluation of the exponent (pow() promotion) agree bit for bit here
    e32 = (-(x * x + y * y).astype(np.float32) / np.float32(450.0)).astype(np.float32)
    e64 = (-(x * x + y * y) / 450.0).astype(np.float32)
    assert np.array_equal(e32, e64)
    g = oracle.gaussian_1d(64, 5.0)
    i = np.arange(64, dtype=np.float64)
    assert np.allclose(g, np.exp(-i * i / 50.0) / math.sqrt(2 * pi32 * 25.0), rtol=2e-5, atol=1e-45)  # exponent rounded to fp32 first


def test_bilateral_small(oracle):
    r = rng(16)
    D = 16
    img = r.integers(-8, 8, (12, 20)).astype(np.float32)
    got = oracle.bilateral(img, 3, 5.0, 10.0, D)
    sp = oracle.gaussian_kernel(3, 10.0)
    col = oracle.gaussian_1d(D, 5.0)
    H, W = img.shape
    want = np.zeros_like(img)
    for y in range(H):
        for x in range(W):
            norm = np.float32(0); res = np.float32(0)
            for dy in range(-3, 4):
                for dx in range(-3, 4):
                    vs = img[min(max(y + dy, 0), H - 1), min(max(x + dx, 0), W - 1)]
                    w = np.float32(sp[dy + 3, dx + 3] * col[int(abs(img[y, x] - vs))])
                    norm = np.float32(norm + w)
                    res = fma32(vs, w, res)
            want[y, x] = np.float32(res / norm)
    assert ulp_diff(got, want).max() <= 1     # the float64 stand-in for fma can double-round
    assert (got == want).mean() > 0.95


# ------------------------------------------------------------------- DIBR
def test_occl_bleed_mask(oracle):
    r = rng(17)
    H, W = 9, 30
    dl = (r.random((H, W)) * 10 - 5).astype(np.float32)
    dr = (r.random((H, W)) * 10 - 5).astype(np.float32)
    ol, orr = oracle.occl(dl, dr)
    wl = np.zeros((H, W), np.uint8); wr = np.zeros((H, W), np.uint8)
    for y in range(H):
        for x in range(W):
            wr[y, min(max(x + int(dl[y, x]), 0), W - 1)] = 1            # trunc toward zero (Q21)
            wl[y, min(max(x + int(-dr[y, x]), 0), W - 1)] = 1
    assert np.array_equal(ol, wl) and np.array_equal(orr, wr)
    b = oracle.bleed(ol, 1)
    want = ol.copy()
    for y in range(H):
        for x in range(W):
            cnt = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    sx, sy = x + dx, y + dy
                    if sx < 0: sx = -sx
                    if sy < 0: sy = -sy
                    if sx > W - 1: sx = W - 1 - dx                       # (Q22)
                    if sy > H - 1: sy = H - 1 - dy
                    cnt += ol[sy, sx] > 0
            if cnt > 2.4:
                want[y, x] = 1
    assert np.array_equal(b, want)
    assert np.array_equal(oracle.occl_to_mask(b), (b == 1).astype(np.float32))


def test_bwarp_merge_dbm(oracle):
    r = rng(18)
    H, W = 6, 50
    L = r.integers(0, 256, (H, W, 3), dtype=np.uint8)
    R = r.integers(0, 256, (H, W, 3), dtype=np.uint8)
    dl = (r.random((H, W)) * 12 - 6).astype(np.float32)
    dr = (r.random((H, W)) * 12 - 6).astype(np.float32)
    ml = (r.random((H, W)) < 0.8).astype(np.float32)
    mr = (r.random((H, W)) < 0.8).astype(np.float32)
    shift = np.float32(1.0 - 3.0 / 7.0)
    got = oracle.bwarp(L, mr, dr, -shift)
    xs = np.arange(W, dtype=np.float32)[None, :]
    fx = fma32(np.float32(-shift), dr, np.broadcast_to(xs, dr.shape))   # one rounding, as the PTX's fma
    sx = np.clip(fx, 0, W - 1).astype(np.int64)
    want = (np.take_along_axis(L, sx[..., None].repeat(3, 2), axis=1).astype(np.float32) * mr[..., None]).astype(np.uint8)
    assert np.array_equal(got, want)
    wr = oracle.bwarp(R, ml, dl, np.float32(1.0 - float(shift)))
    tm = oracle.gaussian_dilate((1 - mr).astype(np.float32), 4, 6.0)
    assert (tm >= (1 - mr)).all() and tm.max() <= 1.0 + 1e-6
    merged = oracle.merge_ab(got, wr, tm)
    want_m = ((1 - tm)[..., None] * got.astype(np.float32)).astype(np.uint8) + (tm[..., None] * wr.astype(np.float32)).astype(np.uint8)
    assert np.array_equal(merged, want_m)
    assert np.array_equal(oracle.dbm(L, R, dl, dr, ml, mr, shift, 4, 6.0), merged)


def test_mux_view_pattern(oracle):
    # constant-colour views make the view-selection pattern visible: R from view r, G from r+1, B from r+2 (Q29)
    V, H, W = 8, 16, 32
    views = [np.full((H, W, 3), 10 * (v + 1), np.uint8) for v in range(V)]
    out = oracle.mux_multiview(views, 18.0, H, W, 2)
    yi = np.float32(np.float64(np.float32(8.0)) / math.tan(float(np.float32(18.0) * np.float32(3.1415926535)) / 180.0) / 3.0)
    inv = np.float32(1.0) / yi
    ri = int(np.round(yi))
    for ty in range(H):
        for tx in range(W):
            yv = np.float32(inv * np.float32(np.float32(ty % ri + 1.0) * np.float32(8.0)))
            rv = (tx * 3 + int(yv)) % 8
            assert tuple(out[ty, tx]) == (10 * ((rv + 2) % 8 + 1), 10 * ((rv + 1) % 8 + 1), 10 * (rv + 1))
    # out size == in size resamples at (tx/W)*W, which is not always tx: still within one grey level of the source
    views = [rng(19 + v).integers(0, 256, (H, W, 3), dtype=np.uint8) for v in range(V)]
    assert oracle.mux_multiview(views, 18.0, H, W, 2).shape == (H, W, 3)
    assert oracle.mux_multiview(views, 18.0, 2 * H - 2, 2 * W, 1).shape == (2 * H - 2, 2 * W, 3)


def test_full_pipeline_runs_and_is_deterministic(oracle, bud_sbs):
    sbs = np.ascontiguousarray(np.concatenate([bud_sbs[100:196, 0:320], bud_sbs[100:196, 640:960]], axis=1))
    a = oracle.adcensus_stm(sbs, 320, 96, 320, D=32, zd=16, want_taps=True, **{k: DEFAULTS[k] for k in
                            ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")})
    b = oracle.adcensus_stm(sbs, 320, 96, 320, D=32, zd=16, **{k: DEFAULTS[k] for k in
                            ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")})
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    taps = a[3]
    assert np.array_equal(taps["views"][0], sbs[:, 320:]) and np.array_equal(taps["views"][7], sbs[:, :320])
    assert set(np.unique(taps["wta_l"])) <= set(np.arange(-16, 16, dtype=np.float32))
    dl, dr = oracle.costvol(sbs[:, :320], sbs[:, 320:], 32, 16)
    assert np.array_equal(dl, taps["wta_l"]) and np.array_equal(dr, taps["wta_r"])
