"""Headless drivers (drivers/): the reference's image_io.cpp / video_io.cpp command lines over the symbols
libs2mv.so exports.  CPU: they build and reject bad command lines without touching the GPU.  GPU: their
files equal what the Python mirror of the same entry points returns."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DEFAULTS, ROOT

DRV = os.path.join(ROOT, "drivers")


REFHDR = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module", params=["compat_header", "reference_headers"])
def drivers(s2mv, request):
    """compat_header: drivers/ built against include/s2mv_compat.h.  reference_headers: the SAME sources built
    against the reference's own headers (image_io.cpp:10-26 / video_io.cpp:11-14 include lists, from
    /root/reference, by oracle/build_ref.sh in the container) and linked against libs2mv.so -- the proof that
    the exported prototypes are the reference's, not a hand copy of them."""
    if request.param == "compat_header":
        subprocess.check_call(["make", "-C", DRV, "-s"])
        return os.path.join(DRV, "s2mv_image"), os.path.join(DRV, "s2mv_video")
    exes = os.path.join(REFHDR, "s2mv_image_refhdr"), os.path.join(REFHDR, "s2mv_video_refhdr")
    if not all(os.path.exists(e) for e in exes):
        pytest.skip("oracle/_ref/s2mv_*_refhdr not built (oracle/build_ref.sh needs /root/reference)")
    return exes


def _write_bmp(path, bgr):
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(bgr[..., ::-1])).save(path, format="BMP")


def _read_bmp(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))[..., ::-1]


def test_drivers_build_and_print_usage(drivers):
    for exe, nargs in zip(drivers, (16, 15)):
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "Usage: ./program" in r.stdout
        r = subprocess.run([exe] + ["1"] * (nargs - 1), capture_output=True, text=True)
        assert r.returncode != 0 and "Usage" in r.stdout
    # unreadable inputs fail before any GPU work
    r = subprocess.run([drivers[0], "/nonexistent/l.bmp", "/nonexistent/r.bmp"] + ["1"] * 14, capture_output=True, text=True)
    assert r.returncode != 0 and "Could not read image files" in r.stdout
    r = subprocess.run([drivers[1], "/nonexistent/v.bgr"] + ["8"] * 14, capture_output=True, text=True)
    assert r.returncode != 0 and "Video cannot be read" in r.stdout


def test_bmp_io_roundtrip_through_the_video_driver_argument_check(drivers, tmp_path):
    # a BMP written by PIL is read by the driver (bottom-up rows, 4-byte row padding: width 3 px) and the
    # parameter check rejects num_views = 1 after reading it, still without a GPU
    img = (np.arange(5 * 6 * 3, dtype=np.uint8).reshape(5, 6, 3) * 7)
    _write_bmp(tmp_path / "f.bmp", img)
    r = subprocess.run([drivers[1], str(tmp_path / "f.bmp"), "1", "18", "3", "5", "8", "4", "10", "30", "20", "6", "17", "9",
                        "20", "0.4"], capture_output=True, text=True, env=dict(os.environ, S2MV_OUT=str(tmp_path)))
    assert r.returncode != 0 and "Input Width (SBS):       6" in r.stdout and "Parameters out of range" in r.stdout


@pytest.mark.gpu
def test_video_driver_equals_adcensus_stm(drivers, s2mv, tmp_path):
    from s2mv_b200_pkg import synth
    H, W, D, zd = 96, 160, 32, 16
    frames = [synth.make_sbs(H, W, 300 + i) for i in range(3)]
    raw = tmp_path / "clip.bgr"
    raw.write_bytes(b"".join(f.tobytes() for f in frames))
    a = DEFAULTS
    args = [str(raw), "8", "18", str(W), str(H), str(D), str(zd), str(a["ad_coeff"]), str(a["census_coeff"]), str(a["ucd"]),
            str(a["lcd"]), str(a["usd"]), str(a["lsd"]), str(a["thresh_s"]), str(a["thresh_h"])]
    env = dict(os.environ, S2MV_OUT=str(tmp_path), S2MV_SBS_COLS=str(2 * W), S2MV_ROWS=str(H))
    r = subprocess.run([drivers[1]] + args, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "3 frame(s)" in r.stdout
    out = np.fromfile(tmp_path / "interlaced.bgr", np.uint8).reshape(3, H, W, 3)
    dl = np.fromfile(tmp_path / "disp_l.f32", np.float32).reshape(3, H, W)
    dr = np.fromfile(tmp_path / "disp_r.f32", np.float32).reshape(3, H, W)
    algo = {k: a[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}
    with s2mv.Pipeline(0, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **algo) as p:
        for i, f in enumerate(frames):
            wl, wr, wo = p.adcensus_stm(f)
            assert np.array_equal(dl[i], wl) and np.array_equal(dr[i], wr) and np.array_equal(out[i], wo)


@pytest.mark.gpu
def test_image_driver_equals_stage_calls(drivers, s2mv, bud_sbs, tmp_path):
    H, W, D, zd = 128, 192, 48, 20
    L = np.ascontiguousarray(bud_sbs[60:60 + H, 200:200 + W])
    R = np.ascontiguousarray(bud_sbs[60:60 + H, 840:840 + W])
    _write_bmp(tmp_path / "l.bmp", L)
    _write_bmp(tmp_path / "r.bmp", R)
    a = DEFAULTS
    args = [str(tmp_path / "l.bmp"), str(tmp_path / "r.bmp"), str(a["ad_coeff"]), str(a["census_coeff"]), str(D), str(zd),
            str(a["ucd"]), str(a["lcd"]), str(a["usd"]), str(a["lsd"]), "8", "18", str(W), str(H), str(a["thresh_s"]),
            str(a["thresh_h"])]
    r = subprocess.run([drivers[0]] + args, capture_output=True, text=True, env=dict(os.environ, S2MV_OUT=str(tmp_path)))
    assert r.returncode == 0, r.stdout + r.stderr
    with s2mv.Pipeline(0) as p:     # the same stage sequence (image_io.cpp:171-292) through the Python mirror
        cl, cr = p.ci_adcensus(L, R, a["ad_coeff"], a["census_coeff"], D, zd)
        xl, al = p.ca_cross(L, cl, a["ucd"], a["lcd"], a["usd"], a["lsd"])
        xr, ar = p.ca_cross(R, cr, a["ucd"], a["lcd"], a["usd"], a["lsd"])
        dl, dr = p.dc_wta(al, zd), p.dc_wta(ar, zd)
        ol, orr = p.dr_dcc(dl, dr)
        dl, _ = p.dr_irv(dl, ol, xl, a["thresh_s"], a["thresh_h"], D, zd, a["usd"], 1, host_variant=True)
        dr, _ = p.dr_irv(dr, orr, xr, a["thresh_s"], a["thresh_h"], D, zd, a["usd"], 1, host_variant=True)
        dl = p.filter_bilateral_1(dl, 7, 7, 7, D)
        dr = p.filter_bilateral_1(dr, 7, 7, 7, D)
    assert np.array_equal(np.fromfile(tmp_path / "disp_l.f32", np.float32).reshape(H, W), dl)
    assert np.array_equal(np.fromfile(tmp_path / "disp_r.f32", np.float32).reshape(H, W), dr)
    # outer views are the inputs (image_io.cpp:270-276), the interlaced frame has the requested size
    assert np.array_equal(_read_bmp(tmp_path / "view_0.bmp"), R) and np.array_equal(_read_bmp(tmp_path / "view_7.bmp"), L)
    assert _read_bmp(tmp_path / "interlaced.bmp").shape == (H, W, 3)
