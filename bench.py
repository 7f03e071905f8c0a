#!/usr/bin/env python
"""bench.py — headline benchmark of the stereo -> multiview frame pipeline.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): frames/s, 1920x1080 side-by-side stereo pair, D=128 ->
two refined disparity maps + the 8-view interlaced frame.  A step = one frame
through the whole adcensus_stm-equivalent path (d_io.cu:7-238).

Workload (BASELINE config 2): the bundled img/fish_1 + img/fish_2 pair
(tests/golden/fish_1_2.npz) upscaled to 1920x1080 with the reference's own
bilinear formula (d_tx_scale.cu:30-52); D=128, zd=64 and BASELINE.md §4's
parameters.  With N GPUs every rank processes its own K frames of the stream
(frame-parallel, no collective on the data path): weak scaling.

  value   frames/s with the input frame already resident in HBM (CUDA events
          around s2mv_process_sbs_device, summed over the K steps, max over
          ranks); L2 is flushed (256 MB write) between steps, untimed.
  e2e     frames/s through the host-buffer C ABI, H2D of the frame and D2H of
          both disparity maps and the interlaced frame inside the timed region
          every step: the asynchronous frame stream (s2mv_stream_submit /
          s2mv_stream_collect, 3 frames in flight: copies overlap kernels), with
          the synchronous adcensus_stm-contract call (s2mv_process_sbs) reported
          beside it as e2e.synchronous_call.
  roofline  for the slowest cost-volume kernel: algorithmic bytes per launch
          (stage-separable model of BASELINE.md §3: 40 B per disparity
          evaluation = 20 V per frame, split 6 V / 4 V / 4 V / 6 V over the four
          kernels) / its mean duration from CUDA events inside the timed steps,
          against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (C restatement of the reference kernels, OpenMP
          over all host cores) on a bounded sample of the same frame.

--impl reference: the reference is CUDA-only; its CPU form is the same C
restatement (oracle/), timed on all host cores, each step a bounded sample
(a row band of the frame, extrapolated to a frame — stated in `sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

H, W, D, ZD, V = 1080, 1920, 128, 64, 8
ALGO = dict(ad_coeff=10.0, census_coeff=30.0, ucd=20.0, lcd=6.0, usd=17, lsd=9, thresh_s=20, thresh_h=0.4)
WORKLOAD = "config2: img/fish_1+fish_2 bilinear-upscaled to 1920x1080, D=128, zd=64, 8 views, angle 18"
METRIC = "1080p D=128 stereo->8-view frames/s"
BAND_ROWS = 135  # bounded CPU sample: 1/8 of the frame


WORKLOAD3 = ("config3: synthetic 1080p stereo stream (generator of stereo-to-multiview-cuda_b200/synth.py, seeds 1000+i, "
             "4 distinct frames per rank cycled), D=128, zd=64, 8 views, angle 18")


def load_frames(workload, rank=0):
    """The frames one rank cycles through: config2 = the one bundled pair, config3 = synthetic stream frames."""
    if workload == "config2":
        return [load_frame()], WORKLOAD
    import s2mv_b200  # noqa: F401
    from s2mv_b200_pkg import synth
    return [synth.make_sbs(H, W, 1000 + 4 * rank + i) for i in range(4)], WORKLOAD3


def load_frame():
    import s2mv_b200  # noqa: F401
    from s2mv_b200_pkg import synth
    sbs = np.load(os.path.join(ROOT, "tests", "golden", "fish_1_2.npz"))["sbs"]
    left = synth.upscale_bilinear(sbs[:, :640], H, W)
    right = synth.upscale_bilinear(sbs[:, 640:], H, W)
    return np.ascontiguousarray(np.concatenate([left, right], axis=1))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_sample(sbs, threads_note=True):
    """Oracle, all host cores, on a BAND_ROWS-row band of the frame (full pipeline)."""
    import oracle_py
    y0 = (H - BAND_ROWS) // 2
    band = np.ascontiguousarray(sbs[y0:y0 + BAND_ROWS])
    t0 = time.perf_counter()
    oracle_py.adcensus_stm(band, W, BAND_ROWS, W, num_views=V, angle=18, D=D, zd=ZD, **ALGO)
    dt = time.perf_counter() - t0
    return dt, oracle_py.max_threads(), (f"rows {y0}..{y0 + BAND_ROWS} of the frame ({W}x{BAND_ROWS}, 1/{H // BAND_ROWS} "
                                         f"frame), full pipeline, extrapolated x{H // BAND_ROWS}")


def run_reference_arm(args):
    """The reference is CUDA-only: its CPU implementation is the C restatement under oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun pins OMP_NUM_THREADS=1 in every worker; this arm is rank 0 alone on the host's cores
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle_py
    oracle_py.build()
    frames, workload = load_frames(args.workload)
    sbs = frames[0]
    for _ in range(args.warmup):
        cpu_sample(sbs)
    t = 0.0
    sample = cores = None
    for _ in range(args.steps):
        dt, cores, sample = cpu_sample(sbs)
        t += dt
    per_frame = (t / args.steps) * (H // BAND_ROWS)
    fps = 1.0 / per_frame
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "bundled fish pair, upscaled (synthetic size)",
        "config": {"workload": workload, "step": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the headline): the bundled pair at 1080p; config3: synthetic 1080p stream")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # the single JSON line must be the only thing on stdout: NCCL prints its version banner there at
    # NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import s2mv_b200
    from s2mv_b200_pkg import sharding

    rank, world, local_rank = sharding.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init_process_group("nccl")
    K, Wm = args.steps, max(args.warmup, 3)

    frames, workload = load_frames(args.workload, rank)
    sbs = frames[0]
    NF = len(frames)
    pipe = s2mv_b200.Pipeline(local_rank, num_rows=H, num_cols=W, num_disp=D, zero_disp=ZD, num_views=V, angle=18, **ALGO)
    stream = torch.cuda.current_stream().cuda_stream
    d_frames = [torch.from_numpy(f).to(dev) for f in frames]
    d_dl = torch.empty((H, W), dtype=torch.float32, device=dev)
    d_dr = torch.empty_like(d_dl)
    d_out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device(i=0):
        pipe.process_device(d_frames[i % NF].data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), stream)

    # ---- device-resident throughput ------------------------------------
    for _ in range(Wm):
        step_device()
    torch.cuda.synchronize()
    pipe.enable_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    stage_ms = {k: 0.0 for k in ("prepare", "costvol", "refine", "dibr")}
    kern_ms = {k: 0.0 for k in ("ci_h1", "v2", "v3", "h4_wta")}
    sampler = ClockSampler(local_rank)
    sharding.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches = 0
    for i in range(K):
        flush.fill_(i & 0xff)                      # L2 flush between timed iterations (not timed)
        ev[i][0].record()
        step_device(i)
        ev[i][1].record()
        ev[i][1].synchronize()
        for k, v in pipe.last_timings().items():
            stage_ms[k] += v
        for k, v in pipe.last_costvol_kernel_timings().items():
            kern_ms[k] += v
        launches += pipe.last_launch_count
    torch.cuda.synchronize()
    sharding.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    pipe.enable_timing(False)
    fps, total_frames, slowest_s = sharding.aggregate_throughput(K, dev_ms / 1e3, dev)

    # ---- end to end through the host-buffer C-ABI call -------------------
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]
    h_dl = torch.empty((H, W), dtype=torch.float32).pin_memory()
    h_dr = torch.empty((H, W), dtype=torch.float32).pin_memory()
    h_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    np_ins, np_dl, np_dr, np_out = [h.numpy() for h in h_frames], h_dl.numpy(), h_dr.numpy(), h_out.numpy()
    for _ in range(Wm):
        pipe.adcensus_stm_into(np_ins[0], np_dl, np_dr, np_out)
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        pipe.adcensus_stm_into(np_ins[i % NF], np_dl, np_dr, np_out)   # synchronous: returns after the D2H copies
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    sharding.barrier()
    e2e_sync_fps, _, _ = sharding.aggregate_throughput(K, e2e_sync_s, dev)
    assert np.array_equal(np_out, d_out.cpu().numpy()), "host and device entry points disagree"

    # ---- the same call with PAGEABLE caller buffers (what an unchanged video_io.cpp passes) ----
    pg_in = [np.array(f) for f in frames]
    pg_dl, pg_dr, pg_out = np.empty((H, W), np.float32), np.empty((H, W), np.float32), np.empty((H, W, 3), np.uint8)
    pageable = {}
    for mode in ("staged", "registered"):
        pipe.set_host_registration(mode == "registered")
        for _ in range(Wm):
            pipe.adcensus_stm_into(pg_in[0], pg_dl, pg_dr, pg_out)
        sharding.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            pipe.adcensus_stm_into(pg_in[i % NF], pg_dl, pg_dr, pg_out)
        dt = time.perf_counter() - t0
        pageable[mode] = sharding.aggregate_throughput(K, dt, dev)[0]
        assert np.array_equal(pg_out, np_out)
    pipe.set_host_registration(False)

    # ---- end to end through the asynchronous frame stream (the video loop) ----
    # every frame: host frame -> the slot's pinned buffer -> H2D -> all kernels -> D2H of both disparity
    # maps and the interlaced frame into pinned host memory; copies of neighbouring frames overlap the kernels
    DEPTH = 3
    pipe.stream_open(DEPTH)

    def stream_run(n):
        got = None
        for i in range(n):
            if pipe.stream_pending == DEPTH:
                got = pipe.stream_collect(copy=False)
            np.copyto(pipe.stream_input_buffer(), np_ins[i % NF])   # stands in for the decoder writing the frame
            pipe.stream_submit(None)
        while pipe.stream_pending:
            got = pipe.stream_collect(copy=False)
        return got

    stream_run(Wm)
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = stream_run(K)
    e2e_s = time.perf_counter() - t0
    sharding.barrier()
    clocks = sampler.stop()
    e2e_fps, _, _ = sharding.aggregate_throughput(K, e2e_s, dev)
    assert np.array_equal(got[2], np_out) and np.array_equal(got[0], np_dl), "stream and synchronous entry points disagree"
    pipe.stream_close()

    if rank != 0:
        return 0

    # ---- roofline of the slowest cost-volume kernel ----------------------
    Vbytes = W * H * D * 4                                # one view's cost volume
    alg = {"ci_h1": 6 * Vbytes, "v2": 4 * Vbytes, "v3": 4 * Vbytes, "h4_wta": 6 * Vbytes}
    dom = max(kern_ms, key=kern_ms.get)
    dom_ms = kern_ms[dom] / K
    peak, peak_src = measured_peak()
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    costvol_ms = (stage_ms["prepare"] + stage_ms["costvol"]) / K
    de = 2.0 * W * H * D
    roofline = {"bound": "hbm", "kernel": {"ci_h1": "k_line<LM_CI_H> (cost init + H pass 1)", "v2": "k_line<LM_V> (V pass 2)",
                                           "v3": "k_line<LM_V> (V pass 3)", "h4_wta": "k_line<LM_H_WTA> (H pass 4 + WTA)"}[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": dom_ms,
                "costvol_leg": {"ms": costvol_ms, "mde_per_s": de / (costvol_ms * 1e-3) / 1e6,
                                "achieved_gbs_40B_per_de": 40.0 * de / (costvol_ms * 1e-3) / 1e9,
                                "frac": 40.0 * de / (costvol_ms * 1e-3) / 1e9 / peak}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        dt, cores, sample = cpu_sample(sbs)
        cpu = {"value": 1.0 / (dt * (H // BAND_ROWS)), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": sample}

    print(json.dumps({
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": 1e3 * slowest_s / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": ("bundled fish pair upscaled to 1080p (no published dataset for this path)"
                                 if args.workload == "config2" else "synthetic stereo stream (no published dataset for this path)"),
        "config": {"workload": workload, "parallelism": f"frame-parallel x{world}, no collectives",
                   "l2": "256 MB flush write between timed steps; per-frame volume traffic >> L2"},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(sbs.nbytes),
                "d2h_bytes_per_step": int(np_dl.nbytes + np_dr.nbytes + np_out.nbytes),
                "api": f"s2mv_stream_submit/collect, {DEPTH} frames in flight (host frame -> pinned slot -> H2D -> kernels -> D2H)",
                "synchronous_call": {"value": e2e_sync_fps, "unit": "frames/s", "api": "s2mv_process_sbs (adcensus_stm contract)",
                                     "pageable_caller_buffers": {"staged_through_pinned": pageable["staged"],
                                                                 "page_locked_in_place": pageable["registered"],
                                                                 "unit": "frames/s"}}},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms": {k: v / K for k, v in stage_ms.items()},
        "costvol_kernel_ms": {k: v / K for k, v in kern_ms.items()},
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
