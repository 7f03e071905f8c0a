"""GPU parity of the whole frame path (the `adcensus_stm` replacement,
d_io.cu:7-238) against the CPU oracle, through the C ABI with HOST buffers,
plus size-independent properties at BASELINE.json's full sizes.

Bars: WTA disparities, outlier labels, voted disparities,
masks, warped views and the interlaced frame bit-exact (the oracle is fed the
GPU's two exponential tables, the only arithmetic a CPU cannot reproduce);
post-bilateral disparities bit-exact as well (same operation order).
"""
import hashlib

import numpy as np
import pytest

from conftest import DEFAULTS

pytestmark = pytest.mark.gpu

ALGO = {k: DEFAULTS[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd", "thresh_s", "thresh_h")}


def run_both(pipe, oracle, sbs, W, D, zd, Ho=None, Wo=None):
    H = sbs.shape[0]
    Ho, Wo = Ho or H, Wo or W
    pipe.configure(num_rows=H, num_cols=W, num_rows_out=Ho, num_cols_out=Wo, num_disp=D, zero_disp=zd,
                   num_views=8, angle=18, **ALGO)
    pipe.enable_taps(True)
    dl, dr, out = pipe.adcensus_stm(sbs)
    taps = pipe.read_taps()
    pipe.enable_taps(False)
    luts = pipe.exp_tables()
    o = oracle.adcensus_stm(sbs, W, Ho, Wo, num_views=8, angle=18, D=D, zd=zd, luts=luts, want_taps=True, **ALGO)
    return (dl, dr, out, taps), o


def assert_frame_equal(got, want):
    dl, dr, out, taps = got
    odl, odr, oout, otaps = want
    for k in ("arms_l", "arms_r", "wta_l", "wta_r", "outliers_l", "outliers_r", "irv_l", "irv_r", "mask_l", "mask_r"):
        assert np.array_equal(taps[k], otaps[k]), k
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    assert np.array_equal(taps["views"], otaps["views"])
    assert np.array_equal(out, oout)


def test_config1_bud_full_frame_bit_exact(pipe, oracle, bud_sbs):
    # BASELINE config 1 (in-domain stand-in): bud_2 + bud_3, 640x384, D=64, zd=32
    got, want = run_both(pipe, oracle, bud_sbs, 640, 64, 32)
    assert_frame_equal(got, want)
    assert pipe.last_launch_count > 0


def test_fish_identical_pair(pipe, oracle, fish_sbs):
    # the bundled fish_1/fish_2 are byte-identical images: zero disparity is a minimum everywhere
    got, want = run_both(pipe, oracle, fish_sbs, 640, 64, 32)
    assert_frame_equal(got, want)


@pytest.mark.parametrize("H,W,D,zd,Ho,Wo", [(70, 200, 20, 7, 70, 200), (45, 331, 48, 24, 64, 400),
                                            (64, 161, 128, 64, 64, 161), (33, 96, 5, 2, 33, 96)])
def test_ragged_shapes_bit_exact(pipe, oracle, bud_sbs, H, W, D, zd, Ho, Wo):
    sbs = np.ascontiguousarray(np.concatenate([bud_sbs[50:50 + H, 100:100 + W], bud_sbs[50:50 + H, 740:740 + W]], 1))
    got, want = run_both(pipe, oracle, sbs, W, D, zd, Ho, Wo)
    assert_frame_equal(got, want)


def test_sbs_wider_than_two_views(pipe, oracle, bud_sbs):
    # num_cols_sbs > 2*num_cols: the extra columns are ignored
    W = 300
    sbs = np.ascontiguousarray(bud_sbs[:64, :640])
    got, want = run_both(pipe, oracle, sbs, W, 32, 16)
    assert_frame_equal(got, want)


def test_d256_multi_chunk_wta(pipe, oracle):
    # num_disp > 128 splits the disparity axis over CTAs; WTA is then a 64-bit atomicMin on (cost, d)
    from s2mv_b200_pkg import synth
    sbs = synth.make_sbs(48, 320, 4000)
    got, want = run_both(pipe, oracle, sbs, 320, 256, 128)
    assert_frame_equal(got, want)


@pytest.mark.parametrize("D,zd", [(256, 128), (200, 90), (300, 0)])
def test_chunk_sequential_volumes_match_resident_volumes(s2mv, oracle, D, zd):
    # num_disp > 128 with ONE 128-disparity chunk of the volumes resident at a time (what 8K D=512 needs on
    # one GPU): the same frame bit for bit as the fully resident layout and as the oracle; the stage entry
    # points that exchange whole volumes refuse to run on such an arena
    from s2mv_b200_pkg import synth
    sbs = synth.make_sbs(40, 336, 4100 + D)
    with s2mv.Pipeline(0) as p:
        p.set_chunk_sequential(1)
        got, want = run_both(p, oracle, sbs, 336, D, zd)
        assert p.chunk_sequential
        seq_bytes = p.arena_bytes
        assert_frame_equal(got, want)
        img = np.ascontiguousarray(sbs[:, :336])
        with pytest.raises(s2mv.S2mvError):
            p.ci_adcensus(img, img, 10.0, 30.0, D, zd)
        p.set_chunk_sequential(0)
        got2, _ = run_both(p, oracle, sbs, 336, D, zd)
        assert not p.chunk_sequential and p.arena_bytes > seq_bytes
        assert_frame_equal(got2, want)


def test_device_entry_points_match_host_entry(pipe, bud_sbs):
    import torch
    H, W = 384, 640
    pipe.configure(num_rows=H, num_cols=W, num_disp=64, zero_disp=32, **ALGO)
    dl, dr, out = pipe.adcensus_stm(bud_sbs)
    d_sbs = torch.from_numpy(bud_sbs).cuda()
    d_dl = torch.empty((H, W), dtype=torch.float32, device="cuda")
    d_dr = torch.empty_like(d_dl)
    d_out = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    pipe.process_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_dl.cpu().numpy(), dl) and np.array_equal(d_dr.cpu().numpy(), dr)
    assert np.array_equal(d_out.cpu().numpy(), out)
    # cost-volume leg alone returns the WTA disparities
    pipe.enable_taps(True)
    pipe.adcensus_stm(bud_sbs)
    taps = pipe.read_taps()
    pipe.enable_taps(False)
    pipe.costvol_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_dl.cpu().numpy(), taps["wta_l"]) and np.array_equal(d_dr.cpu().numpy(), taps["wta_r"])


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _upscaled_1080p(src):
    from s2mv_b200_pkg import synth
    L = synth.upscale_bilinear(src[:, :640], 1080, 1920)
    R = synth.upscale_bilinear(src[:, 640:], 1080, 1920)
    return np.ascontiguousarray(np.concatenate([L, R], axis=1))


def test_config2_1080p_d128_costvol_vs_oracle(pipe, oracle, fish_sbs):
    # BASELINE config 2 (the bench frame): fish pair upscaled to 1920x1080 with the reference's bilinear
    # formula, D=128, zd=64 -- the cost-volume entry point alone, WTA disparities bit-exact at full size.
    # (fish_1 / fish_2 are byte-identical images; the full-frame test below runs the non-degenerate pairs.)
    import torch
    sbs = _upscaled_1080p(fish_sbs)
    pipe.configure(num_rows=1080, num_cols=1920, num_disp=128, zero_disp=64, **ALGO)
    d_sbs = torch.from_numpy(sbs).cuda()
    d_dl = torch.empty((1080, 1920), dtype=torch.float32, device="cuda")
    d_dr = torch.empty_like(d_dl)
    pipe.costvol_device(d_sbs.data_ptr(), 3840, d_dl.data_ptr(), d_dr.data_ptr(), None)
    pipe.synchronize()
    odl, odr = oracle.costvol(sbs[:, :1920], sbs[:, 1920:], 128, 64, luts=pipe.exp_tables(),
                              **{k: ALGO[k] for k in ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd")})
    assert np.array_equal(d_dl.cpu().numpy(), odl)
    assert np.array_equal(d_dr.cpu().numpy(), odr)


@pytest.mark.parametrize("kind", ["fish_upscaled", "bud_upscaled", "synth_seed1000"])
def test_1080p_d128_full_frame_every_tap_vs_oracle(pipe, oracle, fish_sbs, bud_sbs, kind):
    # The headline geometry (1920x1080, D=128, zd=64, 8 views) through the host-buffer call, EVERY output and
    # tap against the oracle bit for bit: arms, WTA, cross-check labels, voted disparities (both views), masks,
    # all eight views, both filtered disparity maps and the interlaced frame -- on the bench frame (fish), on a
    # non-degenerate bundled pair (bud) and on a config-3 stream frame (half of the pixels fail the cross-check).
    from s2mv_b200_pkg import synth
    sbs = {"fish_upscaled": lambda: _upscaled_1080p(fish_sbs), "bud_upscaled": lambda: _upscaled_1080p(bud_sbs),
           "synth_seed1000": lambda: synth.make_sbs(1080, 1920, 1000)}[kind]()
    got, want = run_both(pipe, oracle, sbs, 1920, 128, 64)
    assert_frame_equal(got, want)
    dl, dr, out, taps = got
    # size-independent properties on top of the equality
    assert np.array_equal(taps["views"][0], sbs[:, 1920:]) and np.array_equal(taps["views"][7], sbs[:, :1920])
    for k in ("wta_l", "wta_r", "irv_l", "irv_r"):
        assert taps[k].min() >= -64 and taps[k].max() <= 63 and np.array_equal(taps[k], np.rint(taps[k]))
    assert set(np.unique(taps["outliers_l"])) <= {0, 1, 2} and set(np.unique(taps["mask_l"])) <= {0.0, 1.0}
    assert dl.min() >= -64 and dl.max() <= 63.001
    if kind != "fish_upscaled":
        assert (taps["outliers_l"] != 0).mean() > 0.05      # the refinement stages really have work to do
    again = pipe.adcensus_stm(sbs)
    assert digest(dl, dr, out) == digest(*again)            # deterministic


@pytest.mark.gpu
def test_frame_stream_matches_synchronous_call():
    """s2mv_stream_*: same frames, same order, same bytes as s2mv_process_sbs; error behaviour of a full /
    empty stream."""
    import s2mv_b200
    from s2mv_b200_pkg import synth
    H, W, D, zd = 96, 320, 32, 16
    frames = [synth.make_sbs(H, W, 100 + i) for i in range(7)]
    with s2mv_b200.Pipeline(0, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO) as p:
        want = [p.adcensus_stm(f) for f in frames]
        with pytest.raises(s2mv_b200.S2mvError):
            p.stream_submit(frames[0])          # not opened
        p.stream_open(3)
        with pytest.raises(s2mv_b200.S2mvError):
            p.stream_collect()                  # nothing in flight
        got = []
        import torch
        locked = {i: torch.from_numpy(f).pin_memory() for i, f in enumerate(frames) if i % 3 == 2}
        for i, f in enumerate(frames):
            if p.stream_pending == 3:
                got.append(p.stream_collect())
            if i in locked:
                p.stream_submit(locked[i].numpy())   # page-locked: copied to the device from where it lies
            elif i % 2:
                np.copyto(p.stream_input_buffer(), f)
                p.stream_submit(None)
            else:
                p.stream_submit(f)                   # pageable: staged through the slot's pinned buffer
        assert p.stream_pending == 3
        with pytest.raises(s2mv_b200.S2mvError):
            p.stream_submit(frames[0])          # all slots in flight
        while p.stream_pending:
            got.append(p.stream_collect())
        p.stream_close()
        assert len(got) == len(frames)
        for (dl, dr, out), (wl, wr, wo) in zip(got, want):
            assert np.array_equal(dl, wl) and np.array_equal(dr, wr) and np.array_equal(out, wo)
        # the synchronous call still works after the stream is closed
        dl, dr, out = p.adcensus_stm(frames[0])
        assert np.array_equal(out, want[0][2])


@pytest.mark.parametrize("dense_min,list_votes", [("0", "0"), ("0", "1"), ("2000000000", "0")])
def test_region_voting_dense_and_sparse_paths_agree_with_oracle(s2mv, oracle, bud_sbs, dense_min, list_votes, monkeypatch):
    # the three implementations of the vote (per-pixel span histograms summed along each column with a sliding row
    # window / summed per outlier / per-outlier gather) are forced in turn on the same frames: an outlier-heavy
    # synthetic one and a bundled pair
    from s2mv_b200_pkg import synth
    monkeypatch.setenv("S2MV_IRV_DENSE_MIN", dense_min)
    monkeypatch.setenv("S2MV_IRV_LIST_VOTES", list_votes)
    monkeypatch.setenv("S2MV_IRV_COLW", "1")     # the column walk also on these small frames (it is chosen by image size)
    with s2mv.Pipeline(0) as p:
        got, want = run_both(p, oracle, synth.make_sbs(120, 352, 77), 352, 48, 24)
        assert (want[3]["outliers_l"] != 0).mean() > 0.05
        assert_frame_equal(got, want)
        sbs = np.ascontiguousarray(np.concatenate([bud_sbs[100:228, :320], bud_sbs[100:228, 640:960]], 1))
        got, want = run_both(p, oracle, sbs, 320, 96, 40)
        assert_frame_equal(got, want)


@pytest.mark.parametrize("colw", ["1", "3", "4"])
def test_region_voting_column_walk_on_ragged_shapes(s2mv, oracle, colw, monkeypatch):
    # the column walk with 1, 3 and 4 columns per ticket on widths that are not multiples of 4 (or of the ticket), a
    # height that is not a multiple of the 32-row strips, and three histogram widths (128, 256 and 384 bins per pixel)
    from s2mv_b200_pkg import synth
    monkeypatch.setenv("S2MV_IRV_DENSE_MIN", "0")
    monkeypatch.setenv("S2MV_IRV_COLW", colw)
    with s2mv.Pipeline(0) as p:
        for H, W, D, zd, seed in ((75, 203, 40, 17, 5), (41, 130, 200, 90, 6), (70, 97, 300, 150, 7)):
            got, want = run_both(p, oracle, synth.make_sbs(H, W, seed), W, D, zd)
            assert (want[3]["outliers_l"] != 0).mean() > 0.02
            assert_frame_equal(got, want)


@pytest.mark.parametrize("coop", ["0", "1", "2"])
def test_region_voting_in_one_cooperative_launch_agrees_with_oracle(s2mv, oracle, bud_sbs, coop, monkeypatch):
    # light frames run all voting iterations in ONE cooperative launch (k_irv_sparse_all) once the previous frame's
    # outlier lists were short (1, the default); 2 forces it on every frame -- also the outlier-heavy one --, 0 never
    # uses it.  Same frames, same bar; the repeated light frame is the one that takes the hinted path.
    from s2mv_b200_pkg import synth
    monkeypatch.setenv("S2MV_IRV_COOP", coop)
    with s2mv.Pipeline(0) as p:
        sbs = np.ascontiguousarray(np.concatenate([bud_sbs[100:228, :320], bud_sbs[100:228, 640:960]], 1))
        for _ in range(3):
            got, want = run_both(p, oracle, sbs, 320, 96, 40)
            assert_frame_equal(got, want)
        got, want = run_both(p, oracle, synth.make_sbs(120, 352, 77), 352, 48, 24)   # after light frames: a "scene cut"
        assert_frame_equal(got, want)


def test_cross_check_and_occlusion_marks_as_separate_kernels_agree_with_oracle(s2mv, oracle, bud_sbs, monkeypatch):
    # the frame call forms the cross-check labels and the occlusion masks one image row per block (marks in shared
    # memory: k_dcc_row, k_occl_bleed_mask_row); the memset + kernel-per-step forms they replaced serve rows beyond
    # shared memory and are held to the same bar here
    from s2mv_b200_pkg import synth
    monkeypatch.setenv("S2MV_DCC_SPLIT", "1")
    with s2mv.Pipeline(0) as p:
        got, want = run_both(p, oracle, synth.make_sbs(120, 352, 77), 352, 48, 24)
        assert_frame_equal(got, want)
        sbs = np.ascontiguousarray(np.concatenate([bud_sbs[100:228, :320], bud_sbs[100:228, 640:960]], 1))
        got, want = run_both(p, oracle, sbs, 320, 96, 40)
        assert_frame_equal(got, want)


def test_bilateral_scalar_kernel_agrees_with_oracle(s2mv, oracle, bud_sbs, monkeypatch):
    # the frame call runs the paired (f32x2) bilateral kernel; the one-output-at-a-time kernel it replaced stays
    # selectable and is held to the same bar on the same frame
    monkeypatch.setenv("S2MV_BILATERAL_SCALAR", "1")
    sbs = np.ascontiguousarray(np.concatenate([bud_sbs[100:228, :320], bud_sbs[100:228, 640:960]], 1))
    with s2mv.Pipeline(0) as p:
        got, want = run_both(p, oracle, sbs, 320, 96, 40)
        assert_frame_equal(got, want)


def test_host_registration_of_pageable_buffers(s2mv):
    # pageable numpy buffers: staged by default, page-locked in place with host registration on; the same
    # bytes either way, repeated calls reuse the registration, switching it off releases everything
    from s2mv_b200_pkg import synth
    H, W, D, zd = 96, 320, 32, 16
    frames = [synth.make_sbs(H, W, 900 + i) for i in range(3)]
    with s2mv.Pipeline(0, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18, **ALGO) as p:
        want = [p.adcensus_stm(f) for f in frames]
        p.set_host_registration(True)
        dl = np.empty((H, W), np.float32); dr = np.empty((H, W), np.float32); out = np.empty((H, W, 3), np.uint8)
        buf = np.empty_like(frames[0])
        for rep in range(2):
            for f, w in zip(frames, want):
                buf[...] = f                                   # the caller's one reused frame buffer
                p.adcensus_stm_into(buf, dl, dr, out)
                assert np.array_equal(dl, w[0]) and np.array_equal(dr, w[1]) and np.array_equal(out, w[2])
        keep = [f.copy() for f in frames]                      # more buffers, all alive while registered (the contract)
        outs = [(np.empty((H, W), np.float32), np.empty((H, W), np.float32), np.empty((H, W, 3), np.uint8)) for _ in frames]
        for f, o, w in zip(keep, outs, want):
            p.adcensus_stm_into(f, *o)
            assert all(np.array_equal(a, b) for a, b in zip(o, w))
        p.set_host_registration(False)                         # unregisters everything; staging again
        got = p.adcensus_stm(frames[0])
        assert all(np.array_equal(a, b) for a, b in zip(got, want[0]))
        # mode 2 (what the adcensus_stm shim runs): a buffer is page-locked once the same pointer has arrived in
        # two consecutive calls and released when it stops arriving; the bytes never change
        p.set_host_registration(2)
        for rep in range(3):
            for f, w in zip(frames, want):
                buf[...] = f
                p.adcensus_stm_into(buf, dl, dr, out)          # same four pointers every call
                assert np.array_equal(dl, w[0]) and np.array_equal(dr, w[1]) and np.array_equal(out, w[2])
        other = np.empty_like(out)
        p.adcensus_stm_into(keep[1], dl, dr, other)            # two pointers change: the old ones are let go
        assert np.array_equal(other, want[1][2])
        del buf                                                # ... so their owner may free them
        p.adcensus_stm_into(keep[2], dl, dr, other)
        assert np.array_equal(other, want[2][2])
        with pytest.raises(s2mv.S2mvError):
            p.set_host_registration(3)
        p.set_host_registration(0)


@pytest.mark.parametrize("env", [{}, {"S2MV_NO_VV": "1"}, {"S2MV_LINE_BULK": "1"}, {"S2MV_LINE_V1": "1"}, {"S2MV_L2_CFG": "1"}],
                         ids=["default_fused_vertical", "separate_vertical_passes", "bulk_copies_no_tensor_map", "first_kernel_form",
                              "tile_config_1"])
def test_cost_volume_kernel_forms_agree_with_oracle(s2mv, oracle, env, monkeypatch):
    # The cost volume has several code paths: the persistent pipelined kernels with the two vertical passes fused
    # (default), the same with the vertical passes as two launches, with 512-byte bulk copies instead of tensor
    # copies, the first kernel form (k_line; still what num_disp <= 64 runs), and the second tile geometry.  Each is
    # held to the oracle on frames with several tiles per line, several lines per CTA, a ragged last tile in both
    # directions and more rows than one vertical tile -- D = 128 (one chunk) and D = 200 (two chunks, padded).
    from s2mv_b200_pkg import synth
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with s2mv.Pipeline(0) as p:
        for (H, W, D, zd, seed) in ((150, 700, 128, 64, 31), (70, 330, 200, 90, 32)):
            got, want = run_both(p, oracle, synth.make_sbs(H, W, seed), W, D, zd)
            assert_frame_equal(got, want)
