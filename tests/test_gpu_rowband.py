"""Row-band mode (one frame over several band contexts, SURVEY §8e mode 2) against the single-context
frame: bit-exact disparities and interlaced frame.  The bands here all live on cuda:0 (LocalBands moves the
halos with device copies); tools/rowband_bench.py runs the same schedule one process per GPU over NCCL."""
import numpy as np
import pytest

from conftest import DEFAULTS

pytestmark = pytest.mark.gpu

ALGO = {k: DEFAULTS[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("H,W,D,zd,nbands,fuse", [
    (400, 320, 32, 16, 3, None), (300, 336, 160, 70, 2, None), (531, 200, 64, 32, 4, True), (120, 192, 128, 60, 5, True),
    # bands of >= 2*usd rows with num_disp > 64: the vertical passes run fused, one exchange of 2*usd rows;
    # 40-row bands send the same rows to both neighbours; then the passes kept separate on such a frame, and the
    # library's own choice (None: fused from 40*usd rows per band)
    (300, 336, 128, 60, 3, True), (300, 336, 160, 70, 2, True), (200, 256, 96, 40, 5, True), (300, 336, 128, 60, 3, False),
    (1400, 96, 72, 30, 2, None)])
def test_row_bands_equal_whole_frame(s2mv, H, W, D, zd, nbands, fuse, p2p):
    import torch
    from s2mv_b200_pkg import rowband, synth
    sbs = synth.make_sbs(H, W, 4000 + H)
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
    with s2mv.Pipeline(0, **params) as p:
        wl, wr, wo = p.adcensus_stm(sbs)
    # p2p: halo rows stored by the producing pass straight into the neighbouring band's volume (here: another
    # context on the same GPU), epoch words on the stream; otherwise copied between the passes
    lb = rowband.LocalBands([0] * nbands, p2p=p2p, fuse_vertical=fuse, **params)
    try:
        assert [c.own_rows for c in lb.ctx] == [y1 - y0 for y0, y1 in lb.bands]
        rows = min(c.own_rows for c in lb.ctx)
        fused = D > 64 and ((fuse and rows >= 2 * ALGO["usd"]) or (fuse is None and rows >= 40 * ALGO["usd"]))
        assert all(c.halo_rows == (2 if fused else 1) * ALGO["usd"] for c in lb.ctx)
        d_sbs = torch.from_numpy(sbs).cuda()
        dl, dr, out = lb.process({0: d_sbs}, 2 * W)
        assert np.array_equal(dl.cpu().numpy(), wl)
        assert np.array_equal(dr.cpu().numpy(), wr)
        assert np.array_equal(out.cpu().numpy(), wo)
        if p2p:
            for c in lb.ctx:
                c.status()                                       # every wait met its neighbour
            dl2, dr2, out2 = lb.process({0: d_sbs}, 2 * W)      # a second frame through the same epoch words
            assert np.array_equal(dl2.cpu().numpy(), wl) and np.array_equal(out2.cpu().numpy(), wo)
        # a band context refuses the whole-frame entry points
        with pytest.raises(s2mv.S2mvError):
            lb.ctx[0].pipe.process_device(d_sbs.data_ptr(), 2 * W)
    finally:
        lb.close()


def test_band_configuration_errors(s2mv):
    from s2mv_b200_pkg import rowband
    params = dict(num_rows=200, num_cols=64, num_disp=16, zero_disp=8, **ALGO)
    with pytest.raises(s2mv.S2mvError):
        rowband.RowBand(0, 0, 10, **params)                      # fewer than usd rows
    with pytest.raises(s2mv.S2mvError):
        rowband.RowBand(0, 100, 250, **params)                   # outside the frame
    with pytest.raises(s2mv.S2mvError):
        rowband.RowBand(0, 0, 100, apron=20, **params)           # apron below the vertical reach
    with pytest.raises(s2mv.S2mvError):
        rowband.RowBand(0, 0, 100, num_rows_out=100, **params)   # rescaling output


@pytest.mark.parametrize("p2p", [True, False])
def test_bands_on_two_devices_in_one_process(s2mv, p2p):
    """Needs two GPUs (skipped otherwise): the bands live on different devices of one process, the halo rows
    cross NVLink as peer stores from the producing kernels (p2p) or as device-to-device copies."""
    import torch
    from s2mv_b200_pkg import rowband, synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    H, W, D, zd = 420, 352, 96, 40
    sbs = synth.make_sbs(H, W, 515)
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
    with s2mv.Pipeline(0, **params) as p:
        wl, wr, wo = p.adcensus_stm(sbs)
    lb = rowband.LocalBands([0, 1, 0], p2p=p2p, **params)
    try:
        frames = {d: torch.from_numpy(sbs).to(f"cuda:{d}") for d in (0, 1)}
        for _ in range(2):
            dl, dr, out = lb.process(frames, 2 * W)
            assert np.array_equal(dl.cpu().numpy(), wl) and np.array_equal(dr.cpu().numpy(), wr)
            assert np.array_equal(out.cpu().numpy(), wo)
        if p2p:
            for c in lb.ctx:
                c.status()
    finally:
        lb.close()


def test_lost_neighbour_is_an_error_not_a_frame(s2mv, monkeypatch):
    """A halo wait that runs out (the neighbouring band never ran its pass) must surface as an error from
    s2mv_band_status and from every later call on that band, never as a silently wrong frame."""
    import torch
    from s2mv_b200_pkg import rowband, synth
    monkeypatch.setenv("S2MV_BAND_WAIT_SPINS", "2000")           # ~2 ms instead of ~20 s; read at s2mv_create
    H, W, D, zd = 200, 192, 32, 16
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
    sbs = torch.from_numpy(synth.make_sbs(H, W, 91)).cuda()
    lb = rowband.LocalBands([0, 0], p2p=True, **params)
    try:
        a, b = lb.ctx
        a.prepare(sbs.data_ptr(), 2 * W)
        a.run_pass(1)
        a.run_pass(2)                                            # waits for band b's pass 1, which never runs
        with pytest.raises(s2mv.S2mvError, match="never reached"):
            a.status()
        with pytest.raises(s2mv.S2mvError, match="never reached"):
            a.run_pass(3)                                        # poisoned until reconfigured
        with pytest.raises(s2mv.S2mvError):
            a.finish(None, None, None)
    finally:
        lb.close()


def test_closing_a_band_disconnects_its_neighbours(s2mv):
    """Destroying (or reconfiguring) a band must not leave its neighbours storing halo rows into freed memory:
    the survivors fall back to running without that neighbour."""
    import torch
    from s2mv_b200_pkg import rowband, synth
    H, W, D, zd = 200, 192, 32, 16
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
    sbs = torch.from_numpy(synth.make_sbs(H, W, 92)).cuda()
    lb = rowband.LocalBands([0, 0], p2p=True, **params)
    a, b = lb.ctx
    b.close()                                                    # frees b's volumes; a must let go of them
    try:
        a.prepare(sbs.data_ptr(), 2 * W)
        for k in (1, 2, 3, 4):
            a.run_pass(k)                                        # no peer stores, no waits: nothing to hang on
        a.status()
        torch.cuda.synchronize()
    finally:
        a.close()


def test_dist_band_two_processes_ipc(s2mv):
    """DistBand as it runs on the box: one process per GPU under torchrun, neighbours' volumes mapped through
    CUDA IPC, halo rows as peer stores over NVLink; outputs bit for bit the single-context frame.  Needs two
    GPUs (skipped on the one-GPU test box)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(root, "tools", "rowband_bench.py"), "--height", "432", "--width", "640",
           "--disp", "96", "--steps", "2", "--warmup", "1", "--check", "--transport", "p2p"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["n_gpus"] == 2 and all(res["check"].values()), res
