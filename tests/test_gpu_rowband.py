"""Row-band mode (one frame over several band contexts, SURVEY §8e mode 2) against the single-context
frame: bit-exact disparities and interlaced frame.  The bands here all live on cuda:0 (LocalBands moves the
halos with device copies); tools/rowband_bench.py runs the same schedule one process per GPU over NCCL."""
import numpy as np
import pytest

from conftest import DEFAULTS

pytestmark = pytest.mark.gpu

ALGO = {k: DEFAULTS[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}


@pytest.mark.parametrize("p2p", [True, False])
@pytest.mark.parametrize("H,W,D,zd,nbands", [(400, 320, 32, 16, 3), (300, 336, 160, 70, 2), (531, 200, 64, 32, 4),
                                             (120, 192, 128, 60, 5)])
def test_row_bands_equal_whole_frame(s2mv, H, W, D, zd, nbands, p2p):
    import torch
    from s2mv_b200_pkg import rowband, synth
    sbs = synth.make_sbs(H, W, 4000 + H)
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO)
    with s2mv.Pipeline(0, **params) as p:
        wl, wr, wo = p.adcensus_stm(sbs)
    # p2p: halo rows stored by the producing pass straight into the neighbouring band's volume (here: another
    # context on the same GPU), epoch words on the stream; otherwise copied between the passes
    lb = rowband.LocalBands([0] * nbands, p2p=p2p, **params)
    try:
        assert [c.own_rows for c in lb.ctx] == [y1 - y0 for y0, y1 in lb.bands]
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
