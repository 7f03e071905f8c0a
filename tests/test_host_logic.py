"""CPU tests of the host-side logic: frame sharding across ranks (exercised with
a real world_size-2 gloo process group), the synthetic frame generator, and the
bilinear upscaler used to build BASELINE config 2."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_shard_partitions_the_stream(s2mv):
    from s2mv_b200_pkg.sharding import frame_shard
    for n in (0, 1, 7, 240, 241):
        for world in (1, 2, 3, 8):
            shards = [list(frame_shard(n, r, world)) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(map(len, shards)) - min(map(len, shards)) <= 1
            assert all(s == sorted(s) for s in shards)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_aggregation(s2mv, tmp_path):
    """Two processes on CPU (gloo): each takes its shard of a 9-frame stream, 'processes' it with a fake
    per-frame time, and the aggregate is (all frames) / (slowest rank's time) on both ranks."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import s2mv_b200
        from s2mv_b200_pkg import sharding
        rank, world, _ = sharding.dist_env()
        assert sharding.init_process_group("gloo")
        frames = sharding.frame_shard(9, rank, world)
        seconds = 0.5 * len(frames) * (1 + rank)          # rank 1 is slower per frame
        sharding.barrier()
        fps, total, slowest = sharding.aggregate_throughput(len(frames), seconds)
        print(f"RESULT {{rank}} {{len(frames)}} {{total}} {{slowest}} {{fps}}")
    """))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e
    res = sorted(line.split()[1:] for o, _ in outs for line in o.splitlines() if line.startswith("RESULT"))
    assert [r[1] for r in res] == ["5", "4"]                   # 9 frames -> 5 + 4
    for r in res:
        assert float(r[2]) == 9.0 and float(r[3]) == 4.0 and abs(float(r[4]) - 9.0 / 4.0) < 1e-12


def test_synthetic_frames_are_deterministic_and_piecewise_smooth(s2mv, oracle):
    from s2mv_b200_pkg import synth
    a = synth.make_sbs(96, 160, 1000)
    b = synth.make_sbs(96, 160, 1000)
    c = synth.make_sbs(96, 160, 1001)
    assert a.shape == (96, 320, 3) and a.dtype == np.uint8
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    arms = oracle.cross_arms(np.ascontiguousarray(a[:, :160]), 20.0, 6.0, 17, 9)
    assert 2.0 < arms.mean() < 16.0       # white noise would give ~1, a flat image ~17


def test_upscale_follows_reference_formula(s2mv):
    from s2mv_b200_pkg import synth
    img = np.arange(4 * 6 * 3, dtype=np.uint8).reshape(4, 6, 3) * 3
    assert np.array_equal(synth.upscale_bilinear(img, 4, 6), img)        # identity at equal size
    up = synth.upscale_bilinear(img, 8, 12)
    assert up.shape == (8, 12, 3)
    assert np.array_equal(up[::2, ::2], img)                             # even samples land on source pixels
    assert up[1, 1, 0] == int((img[0, 0, 0] + img[0, 1, 0] + img[1, 0, 0] + img[1, 1, 0]) / 4.0)


def test_row_band_partition_and_schedules(s2mv):
    from s2mv_b200_pkg.rowband import gather_rows, halo_schedule, row_bands
    for H in (64, 1080, 2160, 4321):
        for n in (1, 2, 3, 8):
            bands = row_bands(H, n, min_rows=4)
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            sizes = [y1 - y0 for y0, y1 in bands]
            assert max(sizes) - min(sizes) <= 1
            # every transfer has its mirror image on the peer: what b sends towards side s, the peer receives
            # on the opposite side of its own band
            for b in range(n):
                for peer, send_side, recv_side in halo_schedule(b, n):
                    assert abs(peer - b) == 1 and send_side == recv_side == (0 if peer < b else 1)
                    assert (b, 1 - send_side, 1 - recv_side) in halo_schedule(peer, n)
            # a sub-image's rows are covered exactly once by the bands' own rows
            for b, (y0, y1) in enumerate(bands):
                lo, hi = max(0, y0 - 128), min(H, y1 + 128)
                got = gather_rows(bands, lo, hi - lo)
                assert sum(e - a for _, a, e in got) == hi - lo
                assert [a for _, a, _ in got] == sorted(a for _, a, _ in got)
    import pytest
    with pytest.raises(ValueError):
        row_bands(30, 2, min_rows=17)


@pytest.mark.parametrize("WORLD", [2, 3])
def test_gloo_halo_exchange_and_row_gather(s2mv, tmp_path, WORLD):
    """The distributed transport of the row-band mode on CPU tensors over gloo (world_size 2 and 3): after the
    exchange each rank's halo rows hold the neighbour's edge rows, the row all-gather rebuilds the frame from
    uneven bands, and the disparity-row exchange fills every sub-image's apron (from beyond the nearest
    neighbour too at world_size 3)."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        import s2mv_b200
        from s2mv_b200_pkg import sharding, rowband
        rank, world, _ = sharding.dist_env()
        assert sharding.init_process_group("gloo")
        H, W, usd = 41, 6, 3
        bands = rowband.row_bands(H, world, min_rows=usd)
        y0, y1 = bands[rank]
        frame = torch.arange(H * W, dtype=torch.float32).reshape(H, W)
        # a band's volume rows: own rows + usd halo rows either side (clipped), two views
        vlo, vhi = max(0, y0 - usd), min(H, y1 + usd)
        vol = [torch.full((vhi - vlo, W), -1.0) for _ in range(2)]
        for v in range(2):
            vol[v][y0 - vlo:y1 - vlo] = frame[y0:y1] + 1000 * v
        def send_recv(side, recv):
            if side == 0:
                r0 = (y0 - usd if recv else y0)
                if y0 == 0: return None
            else:
                r0 = (y1 if recv else y1 - usd)
                if y1 == H: return None
            return [vol[v][r0 - vlo:r0 - vlo + usd] for v in range(2)]
        rowband.exchange_halos_dist(send_recv, rank, world, dist)
        for v in range(2):
            assert torch.equal(vol[v], frame[vlo:vhi] + 1000 * v), (rank, v)
        full = rowband.allgather_rows_dist(frame[y0:y1].clone(), bands, dist, torch)
        assert torch.equal(full, frame)
        # disparity rows: every sub-image (own rows + an apron that spans MORE than the neighbouring band here)
        # receives exactly the rows it does not own from the ranks that own them
        apron = 17
        ext = [(max(0, a - apron), min(H, b + apron) - max(0, a - apron)) for a, b in bands]
        send, recv = rowband.disparity_row_plan(bands, ext, rank)
        ly0, rows = ext[rank]
        planes = [torch.full((rows, W), -1.0) for _ in range(2)]
        for v in range(2):
            planes[v][y0 - ly0:y1 - ly0] = frame[y0:y1] + 1000 * v
        rowband.exchange_disparity_rows_dist(planes, ly0, y0, send, recv, dist)
        for v in range(2):
            assert torch.equal(planes[v], frame[ly0:ly0 + rows] + 1000 * v), (rank, v)
        print("RESULT ok", rank)
    """))
    port = _free_port()
    procs = []
    for rank in range(WORLD):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(WORLD), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e
        assert "RESULT ok" in o


def test_cpu_binding_reads_the_gpu_local_cpulist(s2mv, tmp_path, monkeypatch):
    """sharding.bind_to_gpu_cpus: the PCI device's local_cpulist, intersected with the CPUs the process may use;
    anything missing leaves the placement alone."""
    import types
    from s2mv_b200_pkg import sharding
    assert sharding.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.parse_cpulist("") == []
    import torch
    allowed = sorted(os.sched_getaffinity(0))
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    (dev / "local_cpulist").write_text("%d\n" % allowed[0])
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda i: types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0))
    try:
        assert sharding.bind_to_gpu_cpus(0, sysfs=str(tmp_path)) == [allowed[0]]
        assert sorted(os.sched_getaffinity(0)) == [allowed[0]]
    finally:
        os.sched_setaffinity(0, allowed)
    assert sharding.bind_to_gpu_cpus(0, sysfs=str(tmp_path / "nowhere")) is None
    assert sorted(os.sched_getaffinity(0)) == allowed
