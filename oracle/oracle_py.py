"""ctypes binding of the CPU oracle (oracle/s2mv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libs2mv_oracle.so")
CFLAGS = ["-O3", "-mavx2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC"]

AD_LUT = 766
CEN_LUT = 65


def build(force=False):
    src = os.path.join(_HERE, "s2mv_oracle.c")
    hdr = os.path.join(_HERE, "s2mv_oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["gcc", *CFLAGS, src, "-o", _SO, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_hamdist.restype = C.c_int
        _lib.orc_hamdist.argtypes = [C.c_uint64, C.c_uint64]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(x):
    return C.c_float(float(x))


class Taps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "wta_l", "wta_r", "outliers_l", "outliers_r", "irv_l", "irv_r",
        "arms_l", "arms_r", "mask_l", "mask_r", "views", "acost_l", "acost_r")]


def max_threads():
    return lib().orc_max_threads()


def hamdist(a, b):
    return lib().orc_hamdist(int(a), int(b))


def demux_sbs(sbs, W):
    H, Ws, es = sbs.shape
    l = np.zeros((H, W, es), np.uint8)
    r = np.zeros((H, W, es), np.uint8)
    lib().orc_demux_sbs(_p(np.ascontiguousarray(sbs)), _p(l), _p(r), H, Ws, W, es)
    return l, r


def gray(img):
    H, W, es = img.shape
    out = np.zeros((H, W), np.uint8)
    lib().orc_gray(_p(np.ascontiguousarray(img)), _p(out), H, W, es)
    return out


def census(g):
    H, W = g.shape
    out = np.zeros((H, W), np.uint64)
    lib().orc_census(_p(np.ascontiguousarray(g)), _p(out), H, W)
    return out


def ad_cost(img_l, img_r, D, zd):
    H, W, es = img_l.shape
    cl = np.zeros((D, H, W), np.float32)
    cr = np.zeros((D, H, W), np.float32)
    lib().orc_ad_cost(_p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)), _p(cl), _p(cr), D, zd, H, W, es)
    return cl, cr


def census_cost(cen_l, cen_r, D, zd):
    H, W = cen_l.shape
    cl = np.zeros((D, H, W), np.float32)
    cr = np.zeros((D, H, W), np.float32)
    lib().orc_census_cost(_p(np.ascontiguousarray(cen_l)), _p(np.ascontiguousarray(cen_r)), _p(cl), _p(cr), D, zd, H, W)
    return cl, cr


def exp_luts(ad_coeff, census_coeff):
    la = np.zeros(AD_LUT, np.float32)
    lc = np.zeros(CEN_LUT, np.float32)
    lib().orc_exp_luts(_f(ad_coeff), _f(census_coeff), _p(la), _p(lc))
    return la, lc


def ci_adcensus(img_l, img_r, D, zd, ad_coeff, census_coeff, luts=None):
    H, W, es = img_l.shape
    cl = np.zeros((D, H, W), np.float32)
    cr = np.zeros((D, H, W), np.float32)
    la, lc = (None, None) if luts is None else luts
    lib().orc_ci_adcensus(_p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)), _p(cl), _p(cr),
                          _p(la), _p(lc), _f(ad_coeff), _f(census_coeff), D, zd, H, W, es)
    return cl, cr


def cross_arms(img, ucd, lcd, usd, lsd):
    H, W, es = img.shape
    arms = np.zeros((4, H, W), np.uint8)
    lib().orc_cross_arms(_p(np.ascontiguousarray(img)), _p(arms), _f(ucd), _f(lcd), usd, lsd, H, W, es)
    return arms


def ca_pass(cost, arms, direction):
    D, H, W = cost.shape
    out = np.zeros_like(cost)
    lib().orc_ca_pass(_p(np.ascontiguousarray(cost)), _p(out), _p(np.ascontiguousarray(arms)), direction, D, H, W)
    return out


def ca_aggregate(cost, arms):
    D, H, W = cost.shape
    out = np.ascontiguousarray(cost).copy()
    lib().orc_ca_aggregate(_p(out), _p(np.ascontiguousarray(arms)), D, H, W)
    return out


def wta(cost, zd):
    D, H, W = cost.shape
    disp = np.zeros((H, W), np.float32)
    lib().orc_wta(_p(np.ascontiguousarray(cost)), _p(disp), D, zd, H, W)
    return disp


def so(cost, img_own, img_other, view, T, H1, H2, zd, want_cost=False):
    """Scanline optimisation + WTA (PARITY UNPINNED: specification, not a restatement; see s2mv_oracle.c)."""
    D, H, W = cost.shape
    disp = np.zeros((H, W), np.float32)
    out = np.zeros_like(cost) if want_cost else None
    lib().orc_so(_p(np.ascontiguousarray(cost, np.float32)), _p(out), _p(disp), _p(np.ascontiguousarray(img_own)),
                 _p(np.ascontiguousarray(img_other)), int(view), _f(T), _f(H1), _f(H2), D, zd, H, W, 3)
    return (disp, out) if want_cost else disp


def dcc(disp_l, disp_r):
    H, W = disp_l.shape
    ol = np.zeros((H, W), np.uint8)
    orr = np.zeros((H, W), np.uint8)
    lib().orc_dcc(_p(ol), _p(orr), _p(np.ascontiguousarray(disp_l)), _p(np.ascontiguousarray(disp_r)), H, W)
    return ol, orr


def irv(disp, outliers, arms, thresh_s, thresh_h, D, zd, usd, iterations, host_variant=False):
    H, W = disp.shape
    d = np.ascontiguousarray(disp).copy()
    o = np.ascontiguousarray(outliers).copy()
    lib().orc_irv(_p(d), _p(o), _p(np.ascontiguousarray(arms)), thresh_s, _f(thresh_h), H, W, D, zd, usd,
                  iterations, int(host_variant))
    return d, o


def gaussian_kernel(radius, sigma):
    w = 2 * radius + 1
    k = np.zeros((w, w), np.float32)
    lib().orc_gaussian_kernel(_p(k), radius, _f(sigma))
    return k


def gaussian_1d(size, sigma):
    k = np.zeros(size, np.float32)
    lib().orc_gaussian_1d(_p(k), size, _f(sigma))
    return k


def bilateral(img, radius, sigma_color, sigma_spatial, D):
    H, W = img.shape
    out = np.ascontiguousarray(img).copy()
    lib().orc_bilateral(_p(out), radius, _f(sigma_color), _f(sigma_spatial), H, W, D)
    return out


def occl(disp_l, disp_r):
    H, W = disp_l.shape
    ol = np.zeros((H, W), np.uint8)
    orr = np.zeros((H, W), np.uint8)
    lib().orc_occl(_p(ol), _p(orr), _p(np.ascontiguousarray(disp_l)), _p(np.ascontiguousarray(disp_r)), H, W)
    return ol, orr


def bleed(img, radius):
    H, W = img.shape
    out = np.ascontiguousarray(img).copy()
    lib().orc_bleed(_p(out), radius, H, W)
    return out


def occl_to_mask(o):
    H, W = o.shape
    m = np.zeros((H, W), np.float32)
    lib().orc_occl_to_mask(_p(m), _p(np.ascontiguousarray(o)), H, W)
    return m


def gaussian_dilate(img, radius, sigma):
    H, W = img.shape
    out = np.ascontiguousarray(img).copy()
    lib().orc_gaussian_dilate(_p(out), radius, _f(sigma), H, W)
    return out


def bwarp(img_in, mask, disp, shift):
    H, W, es = img_in.shape
    out = np.zeros_like(img_in)
    lib().orc_bwarp(_p(out), _p(np.ascontiguousarray(img_in)), _p(np.ascontiguousarray(mask)),
                    _p(np.ascontiguousarray(disp)), _f(shift), H, W, es)
    return out


def merge_ab(img_b, img_a, mask_a):
    H, W, es = img_b.shape
    out = np.ascontiguousarray(img_b).copy()
    lib().orc_merge_ab(_p(out), _p(np.ascontiguousarray(img_a)), _p(np.ascontiguousarray(mask_a)), H, W, es)
    return out


def dbm(img_l, img_r, disp_l, disp_r, mask_l, mask_r, shift, gauss_radius=10, gauss_sigma=15.0):
    H, W, es = img_l.shape
    out = np.zeros_like(img_l)
    lib().orc_dbm(_p(out), _p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)),
                  _p(np.ascontiguousarray(disp_l)), _p(np.ascontiguousarray(disp_r)),
                  _p(np.ascontiguousarray(mask_l)), _p(np.ascontiguousarray(mask_r)),
                  _f(shift), gauss_radius, _f(gauss_sigma), H, W, es)
    return out


def fwarp(img, disp, shift, want_hits=False):
    """dibr_forward_warp_kernel with the lowest-source-column-wins rule; hits = sources per destination."""
    H, W, es = img.shape
    out = np.zeros_like(img)
    hits = np.zeros((H, W), np.int32)
    lib().orc_fwarp(_p(out), _p(np.ascontiguousarray(img)), _p(np.ascontiguousarray(disp, np.float32)), _f(shift),
                    _p(hits), H, W, es)
    return (out, hits) if want_hits else out


def dibr_dfm(img_l, img_r, disp_l, disp_r, shift):
    H, W, es = img_l.shape
    out = np.zeros_like(img_l)
    lib().orc_dibr_dfm(_p(out), _p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)),
                       _p(np.ascontiguousarray(disp_l, np.float32)), _p(np.ascontiguousarray(disp_r, np.float32)),
                       _f(shift), H, W, es)
    return out


def mux_multiview(views, angle, Ho, Wo, kernel_variant=2):
    views = [np.ascontiguousarray(v) for v in views]
    V = len(views)
    H, W, es = views[0].shape
    out = np.zeros((Ho, Wo, es), np.uint8)
    arr = (C.c_void_p * V)(*[v.ctypes.data for v in views])
    lib().orc_mux_multiview(arr, _p(out), V, _f(angle), H, W, Ho, Wo, es, kernel_variant)
    return out


def costvol(img_l, img_r, D, zd, ad_coeff=10.0, census_coeff=30.0, ucd=20.0, lcd=6.0, usd=17, lsd=9, luts=None):
    H, W, es = img_l.shape
    dl = np.zeros((H, W), np.float32)
    dr = np.zeros((H, W), np.float32)
    la, lc = (None, None) if luts is None else luts
    lib().orc_costvol(_p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)), _p(dl), _p(dr),
                      H, W, es, D, zd, _f(ad_coeff), _f(census_coeff), _f(ucd), _f(lcd), usd, lsd, _p(la), _p(lc))
    return dl, dr


def adcensus_stm(sbs, W, Ho, Wo, num_views=8, angle=18, D=64, zd=32, ad_coeff=10.0, census_coeff=30.0,
                 ucd=20.0, lcd=6.0, usd=17, lsd=9, thresh_s=20, thresh_h=0.4, luts=None, want_taps=False):
    """Returns (disp_l, disp_r, interlaced[, taps dict])."""
    sbs = np.ascontiguousarray(sbs)
    H, Ws, es = sbs.shape
    dl = np.zeros((H, W), np.float32)
    dr = np.zeros((H, W), np.float32)
    out = np.zeros((Ho, Wo, es), np.uint8)
    la, lc = (None, None) if luts is None else luts
    taps = None
    keep = {}
    if want_taps:
        keep = dict(
            wta_l=np.zeros((H, W), np.float32), wta_r=np.zeros((H, W), np.float32),
            outliers_l=np.zeros((H, W), np.uint8), outliers_r=np.zeros((H, W), np.uint8),
            irv_l=np.zeros((H, W), np.float32), irv_r=np.zeros((H, W), np.float32),
            arms_l=np.zeros((4, H, W), np.uint8), arms_r=np.zeros((4, H, W), np.uint8),
            mask_l=np.zeros((H, W), np.float32), mask_r=np.zeros((H, W), np.float32),
            views=np.zeros((num_views, H, W, es), np.uint8),
            acost_l=np.zeros((D, H, W), np.float32), acost_r=np.zeros((D, H, W), np.float32))
        taps = Taps(**{k: v.ctypes.data for k, v in keep.items()})
    lib().orc_adcensus_stm(_p(sbs), _p(dl), _p(dr), _p(out), H, Ws, W, Ho, Wo, es, num_views, int(angle), D, zd,
                           _f(ad_coeff), _f(census_coeff), _f(ucd), _f(lcd), usd, lsd, thresh_s, _f(thresh_h),
                           _p(la), _p(lc), C.byref(taps) if taps is not None else None)
    if want_taps:
        return dl, dr, out, keep
    return dl, dr, out


def scale_bilinear(img, out_rows, out_cols):
    H, W, es = img.shape
    out = np.zeros((out_rows, out_cols, es), np.uint8)
    lib().orc_scale_bilinear(_p(np.ascontiguousarray(img)), _p(out), H, W, out_rows, out_cols, es)
    return out


def disp_scale(disp, out_rows, out_cols, scale):
    H, W = disp.shape
    out = np.zeros((out_rows, out_cols), np.float32)
    lib().orc_disp_scale(_p(out), _p(np.ascontiguousarray(disp, np.float32)), out_rows, out_cols, H, W, _f(scale))
    return out


def adcensus_stm_2(sbs, W, Ho, Wo, Hd, Wd, disp_scale, num_views=8, angle=18, D=64, zd=32, ad_coeff=10.0,
                   census_coeff=30.0, ucd=20.0, lcd=6.0, usd=17, lsd=9, thresh_s=20, thresh_h=0.4, luts=None):
    """adcensus_stm_2 (d_io.cu:240-508) -> (disp_l, disp_r at full resolution, interlaced)."""
    sbs = np.ascontiguousarray(sbs, np.uint8)
    H, Ws, es = sbs.shape
    dl = np.zeros((H, W), np.float32)
    dr = np.zeros((H, W), np.float32)
    out = np.zeros((Ho, Wo, es), np.uint8)
    la, lc = (None, None) if luts is None else luts
    lib().orc_adcensus_stm_2(_p(sbs), _p(dl), _p(dr), _p(out), H, Ws, W, Ho, Wo, Hd, Wd, es, _f(disp_scale),
                             num_views, int(angle), D, zd, _f(ad_coeff), _f(census_coeff), _f(ucd), _f(lcd), usd, lsd,
                             thresh_s, _f(thresh_h), _p(la), _p(lc))
    return dl, dr, out
