"""stereo-to-multiview, B200-native: Python mirror of the reference's operator
interface over the C ABI of include/s2mv.h (libs2mv.so, hand-written sm_100a
CUDA).  Function names, argument order and meaning follow the reference
(`adcensus_stm` d_io.h:32-40 and the host wrappers image_io.cpp:171-292 calls);
arrays are numpy, images are HxWx3 BGR uint8, cost volumes are (D,H,W) float32.

There is NO CPU fallback: importing works anywhere (so the symbol table can be
checked on a CPU box), every compute call raises S2mvError without a B200.
"""
import ctypes as C
import os

import numpy as np

from .build import LIB_PATH, build  # noqa: F401

__all__ = ["S2mvError", "Params", "Pipeline", "lib", "EXPORTED_SYMBOLS"]

# every symbol include/s2mv.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "s2mv_status_string", "s2mv_last_error", "s2mv_default_params", "s2mv_create", "s2mv_destroy",
    "s2mv_configure", "s2mv_arena_bytes", "s2mv_device_sm_count", "s2mv_process_sbs",
    "s2mv_process_sbs_device", "s2mv_costvol_device", "s2mv_synchronize", "s2mv_enable_timing",
    "s2mv_last_timings", "s2mv_last_costvol_kernel_timings", "s2mv_last_launch_count", "s2mv_get_exp_tables", "s2mv_get_ad_terms", "s2mv_enable_taps",
    "s2mv_read_taps", "s2mv_ci_adcensus", "s2mv_gray", "s2mv_census", "s2mv_ci_ad", "s2mv_ci_census",
    "s2mv_ca_cross", "s2mv_dc_wta", "s2mv_dr_dcc", "s2mv_dr_irv", "s2mv_filter_bilateral_1",
    "s2mv_dibr_occl", "s2mv_filter_bleed_1", "s2mv_dibr_occl_to_mask", "s2mv_filter_gaussian_1",
    "s2mv_dibr_dbm", "s2mv_dibr_dfm", "s2mv_mux_multiview",
    "s2mv_stream_open", "s2mv_stream_input_buffer", "s2mv_stream_submit", "s2mv_stream_collect",
    "s2mv_stream_pending", "s2mv_stream_close", "s2mv_set_chunk_sequential", "s2mv_is_chunk_sequential",
    "s2mv_configure_band", "s2mv_configure_band_ex", "s2mv_band_info", "s2mv_band_prepare", "s2mv_band_pass", "s2mv_band_halo",
    "s2mv_band_disp", "s2mv_band_finish", "s2mv_band_connect", "s2mv_band_ipc_export", "s2mv_band_ipc_connect",
    "s2mv_band_status", "s2mv_dc_so", "s2mv_enable_so",
    "s2mv_configure_2", "s2mv_process_sbs_2", "s2mv_process_sbs_2_device", "s2mv_set_host_registration",
]
# the reference's own C++ symbols exported as shims (include/s2mv_compat.h)
COMPAT_SYMBOLS = [
    "_Z12adcensus_stmPhPfS0_S_iiiiiiiiiiffffiiif", "_Z11ci_adcensusPhS_PPfS1_ffiiiii",
    "_Z8ca_crossPhPS_PPfS2_ffiiiiii", "_Z6dc_wtaPPfS_iiii", "_Z6dr_dccPhS_PfS0_ii",
    "_Z6dr_irvPfPhPS0_ifiiiiii", "_Z18filter_bilateral_1Pfiffiii", "_Z9dibr_occlPhS_PfS0_ii",
    "_Z14filter_bleed_1Phiii", "_Z17dibr_occl_to_maskPfS_PhS0_ii", "_Z17filter_gaussian_1Pfifii",
    "_Z8dibr_dbmPhS_S_PfS0_S_S_S0_S0_fiii", "_Z13mux_multiviewPPhS_ifiiiii", "_Z7dc_hsloPPfS_PhS1_fffiiiii",
    "_Z14adcensus_stm_2PhPfS0_S_iiiiiiiifiiiiffffiiif", "_Z8dibr_dfmPhS_S_PfS0_fiii",
]


class S2mvError(RuntimeError):
    pass


class Params(C.Structure):
    """s2mv_params (include/s2mv.h)."""
    _fields_ = [
        ("num_rows", C.c_int), ("num_cols", C.c_int), ("num_rows_out", C.c_int), ("num_cols_out", C.c_int),
        ("elem_sz", C.c_int), ("num_views", C.c_int), ("angle", C.c_int),
        ("num_disp", C.c_int), ("zero_disp", C.c_int),
        ("ad_coeff", C.c_float), ("census_coeff", C.c_float), ("ucd", C.c_float), ("lcd", C.c_float),
        ("usd", C.c_int), ("lsd", C.c_int), ("thresh_s", C.c_int), ("thresh_h", C.c_float),
        ("irv_iterations", C.c_int), ("bilateral_radius", C.c_int),
        ("bilateral_sigma_color", C.c_float), ("bilateral_sigma_spatial", C.c_float),
        ("mask_blur_radius", C.c_int), ("mask_blur_sigma", C.c_float),
    ]


_lib = None


def lib():
    """Load libs2mv.so (built in-tree by build.py).  Fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S2mvError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                            "there is no fallback implementation")
        _lib = C.CDLL(LIB_PATH)
        _lib.s2mv_status_string.restype = C.c_char_p
        _lib.s2mv_last_error.restype = C.c_char_p
        _lib.s2mv_arena_bytes.restype = C.c_size_t
        _lib.s2mv_arena_bytes.argtypes = [C.c_void_p]
        for name in EXPORTED_SYMBOLS:
            fn = getattr(_lib, name)
            if name not in ("s2mv_status_string", "s2mv_last_error", "s2mv_arena_bytes", "s2mv_default_params",
                            "s2mv_destroy"):
                fn.restype = C.c_int
        _lib.s2mv_destroy.restype = None
        _lib.s2mv_destroy.argtypes = [C.c_void_p]
        _lib.s2mv_default_params.restype = None
    return _lib


def _check(status):
    if status != 0:
        L = lib()
        raise S2mvError(f"{L.s2mv_status_string(status).decode()}: {L.s2mv_last_error().decode()}")


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f(x):
    return C.c_float(float(x))


def _ptr_table(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def default_params(**kw):
    p = Params()
    lib().s2mv_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(p, k, v)
    if not p.num_rows_out:
        p.num_rows_out = p.num_rows
    if not p.num_cols_out:
        p.num_cols_out = p.num_cols
    return p


class Pipeline:
    """One context on one GPU: the persistent replacement of what adcensus_stm
    allocates and frees per frame (d_io.cu:43-237)."""

    def __init__(self, device=0, **params):
        self._L = lib()
        self._ctx = C.c_void_p()
        _check(self._L.s2mv_create(C.byref(self._ctx), int(device)))
        self.params = None
        self._stream_cols = None
        if params:
            self.configure(**params)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.s2mv_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration -------------------------------------------------
    def configure(self, **params):
        p = default_params(**params)
        _check(self._L.s2mv_configure(self._ctx, C.byref(p)))
        self.params = p
        return self

    def configure_2(self, num_rows_disp, num_cols_disp, disp_scale, **params):
        """Two-resolution context (adcensus_stm_2): estimation at num_rows_disp x num_cols_disp."""
        p = default_params(**params)
        _check(self._L.s2mv_configure_2(self._ctx, C.byref(p), int(num_rows_disp), int(num_cols_disp), _f(disp_scale)))
        self.params = p
        return self

    def adcensus_stm_2(self, img_sbs):
        """adcensus_stm_2 (d_io.cu:240-508): host arrays in/out, synchronous; full-resolution disparities."""
        p = self.params
        img_sbs = np.ascontiguousarray(img_sbs, np.uint8)
        H, Ws, es = img_sbs.shape
        if H != p.num_rows or es != 3:
            raise S2mvError("frame shape does not match the configured size")
        dl = np.empty((H, p.num_cols), np.float32)
        dr = np.empty((H, p.num_cols), np.float32)
        out = np.empty((p.num_rows_out, p.num_cols_out, 3), np.uint8)
        _check(self._L.s2mv_process_sbs_2(self._ctx, _p(img_sbs), Ws, _p(dl), _p(dr), _p(out)))
        return dl, dr, out

    def set_chunk_sequential(self, mode):
        """-1 auto / 0 never / 1 always (num_disp > 128): one 128-disparity chunk of the volumes resident at a
        time.  Takes effect at the next configure()."""
        _check(self._L.s2mv_set_chunk_sequential(self._ctx, int(mode)))

    @property
    def chunk_sequential(self):
        return bool(self._L.s2mv_is_chunk_sequential(self._ctx))

    @property
    def arena_bytes(self):
        return self._L.s2mv_arena_bytes(self._ctx)

    @property
    def sm_count(self):
        return self._L.s2mv_device_sm_count(self._ctx)

    @property
    def last_launch_count(self):
        return self._L.s2mv_last_launch_count(self._ctx)

    def set_host_registration(self, mode=True):
        """Pageable caller buffers: 0 / False = staged through pinned memory every call, 1 / True = page-locked in
        place the first time they are seen, 2 = auto (page-locked once the same pointer arrives twice in a row,
        released when it stops arriving) -- s2mv.h has the lifetime contract."""
        _check(self._L.s2mv_set_host_registration(self._ctx, int(mode)))

    def set_host_registration_auto(self):
        """What the adcensus_stm shim runs for an unchanged video_io.cpp."""
        self.set_host_registration(2)

    def enable_timing(self, on=True):
        _check(self._L.s2mv_enable_timing(self._ctx, int(on)))

    def enable_taps(self, on=True):
        _check(self._L.s2mv_enable_taps(self._ctx, int(on)))

    def synchronize(self):
        _check(self._L.s2mv_synchronize(self._ctx))

    def last_timings(self):
        ms = (C.c_float * 4)()
        _check(self._L.s2mv_last_timings(self._ctx, ms))
        return dict(zip(("prepare", "costvol", "refine", "dibr"), [float(x) for x in ms]))

    def last_costvol_kernel_timings(self):
        ms = (C.c_float * 4)()
        _check(self._L.s2mv_last_costvol_kernel_timings(self._ctx, ms))
        return dict(zip(("ci_h1", "v2", "v3", "h4_wta"), [float(x) for x in ms]))

    def exp_tables(self, ad_coeff=None, census_coeff=None):
        ad_coeff = self.params.ad_coeff if ad_coeff is None else ad_coeff
        census_coeff = self.params.census_coeff if census_coeff is None else census_coeff
        la = np.zeros(766, np.float32)
        lc = np.zeros(65, np.float32)
        _check(self._L.s2mv_get_exp_tables(self._ctx, _f(ad_coeff), _f(census_coeff), _p(la), _p(lc)))
        return la, lc

    def ad_terms(self, ad_coeff=None):
        """The AD term as the fused kernel computes it in place, for sums 0..765."""
        ad_coeff = self.params.ad_coeff if ad_coeff is None else ad_coeff
        t = np.zeros(766, np.float32)
        _check(self._L.s2mv_get_ad_terms(self._ctx, _f(ad_coeff), _p(t)))
        return t

    # ---- frame entry points ---------------------------------------------
    def adcensus_stm_into(self, img_sbs, disp_l, disp_r, interlaced):
        """adcensus_stm with caller-owned host arrays (d_io.h:32-40): outputs are written in place.
        Pinned arrays (e.g. numpy views of torch pinned tensors) are DMA'd without staging."""
        H, Ws, _ = img_sbs.shape
        _check(self._L.s2mv_process_sbs(self._ctx, _p(img_sbs), Ws, _p(disp_l), _p(disp_r), _p(interlaced)))

    def adcensus_stm(self, img_sbs, want_disp=True, want_interlaced=True):
        """Host arrays in/out, synchronous: the adcensus_stm contract (d_io.cu:7-238)."""
        p = self.params
        img_sbs = np.ascontiguousarray(img_sbs, np.uint8)
        H, Ws, es = img_sbs.shape
        if H != p.num_rows or es != 3:
            raise S2mvError("frame shape does not match the configured size")
        dl = np.empty((H, p.num_cols), np.float32) if want_disp else None
        dr = np.empty((H, p.num_cols), np.float32) if want_disp else None
        out = np.empty((p.num_rows_out, p.num_cols_out, 3), np.uint8) if want_interlaced else None
        _check(self._L.s2mv_process_sbs(self._ctx, _p(img_sbs), Ws, _p(dl), _p(dr), _p(out)))
        return dl, dr, out

    @staticmethod
    def _stream(stream):
        # None -> the context's own stream (NULL in the C ABI); an integer is a cudaStream_t handle, where 0
        # (what torch reports for its default stream) means the legacy default stream = cudaStreamLegacy (0x1)
        if stream is None:
            return C.c_void_p(0)
        return C.c_void_p(int(stream) or 1)

    def process_device(self, d_sbs_ptr, num_cols_sbs, d_disp_l=0, d_disp_r=0, d_interlaced=0, stream=None):
        """Raw device pointers (ints, e.g. torch.Tensor.data_ptr()), asynchronous on `stream`."""
        _check(self._L.s2mv_process_sbs_device(self._ctx, C.c_void_p(d_sbs_ptr), int(num_cols_sbs),
                                               C.c_void_p(d_disp_l), C.c_void_p(d_disp_r),
                                               C.c_void_p(d_interlaced), self._stream(stream)))

    def costvol_device(self, d_sbs_ptr, num_cols_sbs, d_disp_l=0, d_disp_r=0, stream=None):
        _check(self._L.s2mv_costvol_device(self._ctx, C.c_void_p(d_sbs_ptr), int(num_cols_sbs),
                                           C.c_void_p(d_disp_l), C.c_void_p(d_disp_r), self._stream(stream)))

    # ---- asynchronous frame stream (video loop) ---------------------------
    def stream_open(self, depth=3, num_cols_sbs=None):
        """Open `depth` in-flight slots; frames then go submit() ... collect() in order."""
        ncs = 2 * self.params.num_cols if num_cols_sbs is None else int(num_cols_sbs)
        _check(self._L.s2mv_stream_open(self._ctx, int(depth), ncs))
        self._stream_cols = ncs

    def stream_close(self):
        _check(self._L.s2mv_stream_close(self._ctx))

    @property
    def stream_pending(self):
        return self._L.s2mv_stream_pending(self._ctx)

    def stream_input_buffer(self):
        """numpy view of the pinned buffer the next submit() will upload (decode straight into it)."""
        ptr = C.POINTER(C.c_uint8)()
        _check(self._L.s2mv_stream_input_buffer(self._ctx, C.byref(ptr)))
        p = self.params
        return np.ctypeslib.as_array(ptr, shape=(p.num_rows, self._stream_cols, 3))

    def stream_submit(self, img_sbs=None):
        """Enqueue one frame (None: the buffer from stream_input_buffer() is already filled)."""
        if img_sbs is not None:
            img_sbs = np.ascontiguousarray(img_sbs, np.uint8)
            if self._stream_cols is not None and img_sbs.shape != (self.params.num_rows, self._stream_cols, 3):
                raise S2mvError("frame shape does not match the opened stream")
        _check(self._L.s2mv_stream_submit(self._ctx, _p(img_sbs)))

    def stream_collect(self, copy=True):
        """Oldest in-flight frame -> (disp_l, disp_r, interlaced).  copy=False returns views of the slot's
        pinned buffers (valid until `depth` further submits)."""
        p = self.params
        H, W = p.num_rows, p.num_cols
        pl, pr, po = C.POINTER(C.c_float)(), C.POINTER(C.c_float)(), C.POINTER(C.c_uint8)()
        _check(self._L.s2mv_stream_collect(self._ctx, None, None, None, C.byref(pl), C.byref(pr), C.byref(po)))
        dl = np.ctypeslib.as_array(pl, shape=(H, W))
        dr = np.ctypeslib.as_array(pr, shape=(H, W))
        out = np.ctypeslib.as_array(po, shape=(p.num_rows_out, p.num_cols_out, 3))
        return (dl.copy(), dr.copy(), out.copy()) if copy else (dl, dr, out)

    def read_taps(self):
        p = self.params
        H, W, V = p.num_rows, p.num_cols, p.num_views
        t = dict(
            wta_l=np.empty((H, W), np.float32), wta_r=np.empty((H, W), np.float32),
            outliers_l=np.empty((H, W), np.uint8), outliers_r=np.empty((H, W), np.uint8),
            irv_l=np.empty((H, W), np.float32), irv_r=np.empty((H, W), np.float32),
            arms_l=np.empty((4, H, W), np.uint8), arms_r=np.empty((4, H, W), np.uint8),
            mask_l=np.empty((H, W), np.float32), mask_r=np.empty((H, W), np.float32),
            views=np.empty((V, H, W, 3), np.uint8))
        _check(self._L.s2mv_read_taps(self._ctx, *[_p(t[k]) for k in (
            "wta_l", "wta_r", "outliers_l", "outliers_r", "irv_l", "irv_r", "arms_l", "arms_r",
            "mask_l", "mask_r", "views")]))
        return t

    # ---- per-stage operators (host arrays), reference names --------------
    def _cost_tables(self, D, H, W):
        vol = np.empty((D, H, W), np.float32)
        return vol, _ptr_table([vol[d] for d in range(D)])

    def ci_adcensus(self, img_l, img_r, ad_coeff, census_coeff, num_disp, zero_disp):
        return self._ci(self._L.s2mv_ci_adcensus, img_l, img_r, num_disp, zero_disp, (_f(ad_coeff), _f(census_coeff)))

    def ci_ad(self, img_l, img_r, num_disp, zero_disp):
        return self._ci(self._L.s2mv_ci_ad, img_l, img_r, num_disp, zero_disp, ())

    def ci_census(self, img_l, img_r, num_disp, zero_disp):
        return self._ci(self._L.s2mv_ci_census, img_l, img_r, num_disp, zero_disp, ())

    def _ci(self, fn, img_l, img_r, D, zd, extra):
        img_l = np.ascontiguousarray(img_l, np.uint8)
        img_r = np.ascontiguousarray(img_r, np.uint8)
        H, W, es = img_l.shape
        cl, tl = self._cost_tables(D, H, W)
        cr, tr = self._cost_tables(D, H, W)
        _check(fn(self._ctx, _p(img_l), _p(img_r), tl, tr, *extra, D, zd, H, W, es))
        return cl, cr

    def gray(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        H, W, es = img.shape
        out = np.empty((H, W), np.uint8)
        _check(self._L.s2mv_gray(self._ctx, _p(img), _p(out), H, W, es))
        return out

    def census(self, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        H, W = gray.shape
        out = np.empty((H, W), np.uint64)
        _check(self._L.s2mv_census(self._ctx, _p(gray), _p(out), H, W))
        return out

    def ca_cross(self, img, cost, ucd, lcd, usd, lsd):
        """-> (arms (4,H,W) uint8, aggregated cost (D,H,W))"""
        img = np.ascontiguousarray(img, np.uint8)
        cost = np.ascontiguousarray(cost, np.float32)
        D, H, W = cost.shape
        arms = np.empty((4, H, W), np.uint8)
        acost, ta = self._cost_tables(D, H, W)
        tc = _ptr_table([cost[d] for d in range(D)])
        tarms = _ptr_table([arms[k] for k in range(4)])
        _check(self._L.s2mv_ca_cross(self._ctx, _p(img), tarms, tc, ta, _f(ucd), _f(lcd), usd, lsd, D, H, W, 3))
        return arms, acost

    def dc_wta(self, cost, zero_disp):
        cost = np.ascontiguousarray(cost, np.float32)
        D, H, W = cost.shape
        disp = np.empty((H, W), np.float32)
        _check(self._L.s2mv_dc_wta(self._ctx, _ptr_table([cost[d] for d in range(D)]), _p(disp), D, zero_disp, H, W))
        return disp

    def enable_so(self, on=True, T=15.0, H1=1.0, H2=3.0):
        """Frame path: scanline optimisation between aggregation and WTA (off by default)."""
        _check(self._L.s2mv_enable_so(self._ctx, int(on), _f(T), _f(H1), _f(H2)))

    def dc_so(self, cost, img_own, img_other, view, T, H1, H2, zero_disp, want_cost=False):
        """Scanline optimisation + WTA of one view's aggregated cost (the reference's unfinished dc_hslo)."""
        cost = np.ascontiguousarray(cost, np.float32)
        D, H, W = cost.shape
        disp = np.empty((H, W), np.float32)
        out, tout = self._cost_tables(D, H, W) if want_cost else (None, None)
        _check(self._L.s2mv_dc_so(self._ctx, _ptr_table([cost[d] for d in range(D)]), _p(disp), tout,
                                  _p(np.ascontiguousarray(img_own, np.uint8)), _p(np.ascontiguousarray(img_other, np.uint8)),
                                  int(view), _f(T), _f(H1), _f(H2), D, zero_disp, H, W, 3))
        return (disp, out) if want_cost else disp

    def dr_dcc(self, disp_l, disp_r):
        disp_l = np.ascontiguousarray(disp_l, np.float32)
        disp_r = np.ascontiguousarray(disp_r, np.float32)
        H, W = disp_l.shape
        ol = np.empty((H, W), np.uint8)
        orr = np.empty((H, W), np.uint8)
        _check(self._L.s2mv_dr_dcc(self._ctx, _p(ol), _p(orr), _p(disp_l), _p(disp_r), H, W))
        return ol, orr

    def dr_irv(self, disp, outliers, arms, thresh_s, thresh_h, num_disp, zero_disp, usd, iterations,
               host_variant=False):
        d = np.ascontiguousarray(disp, np.float32).copy()
        o = np.ascontiguousarray(outliers, np.uint8).copy()
        arms = np.ascontiguousarray(arms, np.uint8)
        H, W = d.shape
        _check(self._L.s2mv_dr_irv(self._ctx, _p(d), _p(o), _ptr_table([arms[k] for k in range(4)]), thresh_s,
                                   _f(thresh_h), H, W, num_disp, zero_disp, usd, iterations, int(host_variant)))
        return d, o

    def filter_bilateral_1(self, img, radius, sigma_color, sigma_spatial, num_disp):
        out = np.ascontiguousarray(img, np.float32).copy()
        H, W = out.shape
        _check(self._L.s2mv_filter_bilateral_1(self._ctx, _p(out), radius, _f(sigma_color), _f(sigma_spatial),
                                               H, W, num_disp))
        return out

    def dibr_occl(self, disp_l, disp_r):
        disp_l = np.ascontiguousarray(disp_l, np.float32)
        disp_r = np.ascontiguousarray(disp_r, np.float32)
        H, W = disp_l.shape
        ol = np.empty((H, W), np.uint8)
        orr = np.empty((H, W), np.uint8)
        _check(self._L.s2mv_dibr_occl(self._ctx, _p(ol), _p(orr), _p(disp_l), _p(disp_r), H, W))
        return ol, orr

    def filter_bleed_1(self, img, radius):
        out = np.ascontiguousarray(img, np.uint8).copy()
        H, W = out.shape
        _check(self._L.s2mv_filter_bleed_1(self._ctx, _p(out), radius, H, W))
        return out

    def dibr_occl_to_mask(self, occl_l, occl_r):
        occl_l = np.ascontiguousarray(occl_l, np.uint8)
        occl_r = np.ascontiguousarray(occl_r, np.uint8)
        H, W = occl_l.shape
        ml = np.empty((H, W), np.float32)
        mr = np.empty((H, W), np.float32)
        _check(self._L.s2mv_dibr_occl_to_mask(self._ctx, _p(ml), _p(mr), _p(occl_l), _p(occl_r), H, W))
        return ml, mr

    def filter_gaussian_1(self, img, radius, sigma_spatial):
        out = np.ascontiguousarray(img, np.float32).copy()
        H, W = out.shape
        _check(self._L.s2mv_filter_gaussian_1(self._ctx, _p(out), radius, _f(sigma_spatial), H, W))
        return out

    def dibr_dbm(self, img_l, img_r, disp_l, disp_r, mask_l, mask_r, shift, blur_radius=10, blur_sigma=15.0):
        img_l = np.ascontiguousarray(img_l, np.uint8)
        img_r = np.ascontiguousarray(img_r, np.uint8)
        H, W, es = img_l.shape
        out = np.empty((H, W, 3), np.uint8)
        args = [np.ascontiguousarray(a, np.float32) for a in (disp_l, disp_r, mask_l, mask_r)]
        _check(self._L.s2mv_dibr_dbm(self._ctx, _p(out), _p(img_l), _p(img_r), *[_p(a) for a in args], _f(shift),
                                     blur_radius, _f(blur_sigma), H, W, es))
        return out

    def dibr_dfm(self, img_l, img_r, disp_l, disp_r, shift):
        """Forward-warp view synthesis (d_dibr_fwarp.cu:97-197); colliding sources: lowest source column wins."""
        H, W, es = img_l.shape
        out = np.zeros((H, W, es), np.uint8)
        args = [np.ascontiguousarray(a, np.float32) for a in (disp_l, disp_r)]
        _check(self._L.s2mv_dibr_dfm(self._ctx, _p(out), _p(np.ascontiguousarray(img_l)), _p(np.ascontiguousarray(img_r)),
                                     *[_p(a) for a in args], _f(shift), H, W, es))
        return out

    def mux_multiview(self, views, angle, num_rows_out, num_cols_out, kernel_variant=0):
        views = [np.ascontiguousarray(v, np.uint8) for v in views]
        H, W, es = views[0].shape
        out = np.empty((num_rows_out, num_cols_out, 3), np.uint8)
        _check(self._L.s2mv_mux_multiview(self._ctx, _ptr_table(views), _p(out), len(views), _f(angle), H, W,
                                          num_rows_out, num_cols_out, es, kernel_variant))
        return out
