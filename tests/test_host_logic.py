"""CPU tests of the host-side logic: frame sharding across ranks (exercised with
a real world_size-2 gloo process group), the synthetic frame generator, and the
bilinear upscaler used to build BASELINE config 2."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

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
