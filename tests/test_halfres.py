"""adcensus_stm_2 (d_io.cu:240-508): estimation at reduced resolution, DIBR at full resolution.
CPU: the oracle's scaling kernels.  GPU: product vs oracle bit for bit, and vs the reference's own
adcensus_stm_2 (oracle/_ref, built from /root/reference for sm_100) at the in-domain shape."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import DEFAULTS, ROOT

ALGO = {k: DEFAULTS[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}


def test_oracle_scaling_kernels(s2mv, oracle):
    from s2mv_b200_pkg import synth
    img = np.ascontiguousarray(synth.make_sbs(60, 90, 9)[:, :90])
    # identity size: every sample lands on a source pixel ... except where (t / n) * n rounds below t
    same = oracle.scale_bilinear(img, 60, 90)
    assert (same == img).mean() > 0.9
    down = oracle.scale_bilinear(img, 30, 45)
    assert down.shape == (30, 45, 3)
    ref = synth.upscale_bilinear(img, 30, 45).astype(int)              # same formula in numpy, without the fused multiply-adds
    assert np.abs(down.astype(int) - ref).max() <= 1 and (down == ref).mean() > 0.98
    d = np.arange(12, dtype=np.float32).reshape(3, 4)
    up = oracle.disp_scale(d, 6, 8, 2.0)
    assert up.shape == (6, 8) and up[0, 0] == 0.0 and np.isclose(up[0, 2], 2.0 * d[0, 1])
    assert np.all(np.diff(up, axis=1) >= 0) and np.all(np.diff(up, axis=0) >= 0)   # monotone ramp stays monotone


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,Hd,Wd,scale,D,zd", [(96, 160, 48, 80, 0.5, 16, 8), (120, 200, 50, 77, 0.4, 24, 12),
                                                 (64, 128, 64, 128, 1.0, 32, 16)])
def test_product_equals_oracle(s2mv, oracle, H, W, Hd, Wd, scale, D, zd):
    from s2mv_b200_pkg import synth
    sbs = synth.make_sbs(H, W, 600 + H)
    with s2mv.Pipeline(0) as p:
        p.configure_2(Hd, Wd, scale, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
        dl, dr, out = p.adcensus_stm_2(sbs)
        luts = p.exp_tables()
        with pytest.raises(s2mv.S2mvError):
            p.adcensus_stm(sbs)                      # a two-resolution context refuses the one-resolution call
        odl, odr, oout = oracle.adcensus_stm_2(sbs, W, H, W, Hd, Wd, scale, D=D, zd=zd, luts=luts, **ALGO)
        assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
        assert np.array_equal(out, oout)
        # back to a one-resolution context on the same handle
        p.configure(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
        a = p.adcensus_stm(sbs)
        b = oracle.adcensus_stm(sbs, W, H, W, D=D, zd=zd, luts=luts, **ALGO)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.gpu
def test_product_equals_reference_adcensus_stm_2(s2mv, oracle):
    """The reference's own adcensus_stm_2 (one-barrier build, Q15) on a 640x960 synthetic pair estimated at
    320x480: a shape inside the reference's validity domain at BOTH resolutions (W % 160 == 0, H % 32 == 0,
    D <= 65, and H % 30 == 0 so that its bilateral filter fills every tile row -- at 192 low-resolution rows the
    unfilled rows of Q19 index its colour table out of bounds and the reference aborts)."""
    from s2mv_b200_pkg import synth
    so = os.path.join(ROOT, "oracle", "_ref", "libs2mv_ref_q15.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    ref = C.CDLL(so)
    fn = getattr(ref, "_Z14adcensus_stm_2PhPfS0_S_iiiiiiiifiiiiffffiiif")
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int] * 8 + [C.c_float] + [C.c_int] * 4 + [C.c_float] * 4 + [C.c_int] * 3 + [C.c_float]
    H, W, Hd, Wd, D, zd = 960, 640, 480, 320, 32, 16
    sbs = synth.make_sbs(H, W, 4242)
    rdl = np.zeros((H, W), np.float32); rdr = np.zeros((H, W), np.float32); rout = np.zeros((H, W, 3), np.uint8)
    a = DEFAULTS
    fn(sbs.ctypes.data, rdl.ctypes.data, rdr.ctypes.data, rout.ctypes.data, H, 2 * W, W, H, W, Hd, Wd, 3, 0.5, 8, 18, D, zd,
       a["ad_coeff"], a["census_coeff"], a["ucd"], a["lcd"], a["usd"], a["lsd"], a["thresh_s"], a["thresh_h"])
    with s2mv.Pipeline(0) as p:
        p.configure_2(Hd, Wd, 0.5, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
        dl, dr, out = p.adcensus_stm_2(sbs)
    assert np.array_equal(dl, rdl) and np.array_equal(dr, rdr)
    assert np.array_equal(out, rout)
    dst = os.environ.get("S2MV_WRITE_GOLDEN_STM2")
    if dst:     # the REFERENCE's outputs, as hashes, for the CPU test below
        np.savez(dst, params=np.array([H, W, Hd, Wd, D, zd, 4242], np.int32), sha_disp_l=_sha(rdl), sha_disp_r=_sha(rdr),
                 sha_interlaced=_sha(rout), disp_l_row300=rdl[300], interlaced_row300=rout[300])


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_matches_reference_stm_2_golden(s2mv, oracle):
    """CPU: the oracle's adcensus_stm_2 against hashes of the reference's own outputs (written on the GPU box by the
    test above with S2MV_WRITE_GOLDEN_STM2), fed the GPU's exponential tables stored in ref_golden.npz."""
    from conftest import GOLDEN
    from s2mv_b200_pkg import synth
    path, tables = os.path.join(GOLDEN, "ref_golden_stm2.npz"), os.path.join(GOLDEN, "ref_golden.npz")
    if not (os.path.exists(path) and os.path.exists(tables)):
        pytest.skip("golden not generated yet")
    g, t = np.load(path), np.load(tables)
    H, W, Hd, Wd, D, zd, seed = [int(v) for v in g["params"]]
    sbs = synth.make_sbs(H, W, seed)
    dl, dr, out = oracle.adcensus_stm_2(sbs, W, H, W, Hd, Wd, 0.5, D=D, zd=zd, luts=(t["lut_ad"], t["lut_cen"]), **ALGO)
    assert np.array_equal(dl[300], g["disp_l_row300"]) and np.array_equal(out[300], g["interlaced_row300"])
    assert _sha(dl) == str(g["sha_disp_l"]) and _sha(dr) == str(g["sha_disp_r"]) and _sha(out) == str(g["sha_interlaced"])
