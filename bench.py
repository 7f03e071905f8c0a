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
  roofline  for the slowest cost-volume kernel: the bytes its algorithm must
          move per launch (2V / 4V / 4V / 2V for the four fused kernels; ncu's
          DRAM bytes agree to ~1 %) / its mean duration from CUDA events inside
          the timed steps, against the measured HBM copy bandwidth in
          MEASURED_PEAKS.json; `bound` from the ncu capture (hbm | issue); all
          four kernels under roofline.kernels; the stage-separable 40 B/DE
          model only as a separately named throughput.
  cpu_baseline  the CPU oracle (C restatement of the reference kernels, OpenMP
          over all host cores) on one whole frame.

  extra   other workloads through the same code (never the headline): config3
          (synthetic stream frame, half of the pixels fail the cross-check),
          bud_1080p (a non-degenerate bundled pair -- fish_1/fish_2 are
          byte-identical images); reference_gpu = the reference's OWN kernels
          timed on this box (oracle/ref_gpu_time.py, separate process); with
          N > 1: rowband = one 3840x2160 D=256 frame in N row bands.

--impl reference: the reference is CUDA-only; its CPU form is the same C
restatement (oracle/), timed on all host cores, one WHOLE frame per step, same
`config` as the main arm.
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


def config_dict(workload, world):
    """`config` of the JSON line -- the SAME dict for the main arm and for --impl reference."""
    return {"workload": workload, "parallelism": f"frame-parallel x{world}, no collectives",
            "l2": "256 MB flush write between timed steps; per-frame volume traffic >> L2"}


def cpu_frame(sbs):
    """The CPU oracle (C restatement of the reference kernels, OpenMP over all host cores) on ONE WHOLE frame:
    seconds, threads used."""
    import oracle_py
    t0 = time.perf_counter()
    oracle_py.adcensus_stm(sbs, W, H, W, num_views=V, angle=18, D=D, zd=ZD, **ALGO)
    return time.perf_counter() - t0, oracle_py.max_threads()


def cpu_band_x8(sbs):
    """Round 1's bounded sample (a BAND_ROWS-row band, x8), kept only as a cross-check of the whole-frame time."""
    import oracle_py
    y0 = (H - BAND_ROWS) // 2
    band = np.ascontiguousarray(sbs[y0:y0 + BAND_ROWS])
    t0 = time.perf_counter()
    oracle_py.adcensus_stm(band, W, BAND_ROWS, W, num_views=V, angle=18, D=D, zd=ZD, **ALGO)
    return (time.perf_counter() - t0) * (H // BAND_ROWS)


CPU_SAMPLE = "one whole 1920x1080 frame per step, full pipeline (disparities + 8-view interlaced frame), nothing extrapolated"


def run_reference_arm(args):
    """The reference is CUDA-only: its CPU implementation is the C restatement under oracle/, all host cores,
    one WHOLE frame of the main arm's workload per step (same `config`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun pins OMP_NUM_THREADS=1 in every worker; this arm is rank 0 alone on the host's cores
    if "TORCHELASTIC_RUN_ID" in os.environ or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle_py
    oracle_py.build()
    frames, workload = load_frames(args.workload)
    for i in range(args.warmup):
        cpu_frame(frames[i % len(frames)])
    t, cores = 0.0, None
    for i in range(args.steps):
        dt, cores = cpu_frame(frames[i % len(frames)])
        t += dt
    fps = args.steps / t
    band = cpu_band_x8(frames[0])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": DATA[args.workload],
        "config": config_dict(workload, args.gpus),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE,
                         "band_x8_estimate_frames_per_s": 1.0 / band},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


DATA = {"config2": "bundled fish pair upscaled to 1080p (no published dataset for this path)",
        "config3": "synthetic stereo stream (no published dataset for this path)"}
KERNEL_NAMES = {"ci_h1": "k_line2<LM_CI_H> (cost init + H pass 1)", "v2": "k_line2<LM_V> (V pass 2)",
                "v3": "k_line2<LM_V> (V pass 3)", "v2_v3_fused": "k_line_vv (V passes 2 and 3 in one kernel)",
                "h4_wta": "k_line2<LM_H_WTA> (H pass 4 + WTA)"}


def fold_fused(kern_ms):
    """The C ABI reports four intervals; with the two vertical passes fused the third is empty."""
    if kern_ms.get("v3", 1.0) < 0.02 * max(kern_ms.get("v2", 0.0), 1e-9):
        return {"ci_h1": kern_ms["ci_h1"], "v2_v3_fused": kern_ms["v2"] + kern_ms["v3"], "h4_wta": kern_ms["h4_wta"]}
    return dict(kern_ms)


class Rig:
    """One context + device/pinned buffers for 1080p D=128 frames; the three ways a frame is timed."""

    def __init__(self, torch, s2mv_b200, sharding, local_rank):
        self.torch, self.sharding = torch, sharding
        self.dev = torch.device("cuda", local_rank)
        self.pipe = s2mv_b200.Pipeline(local_rank, num_rows=H, num_cols=W, num_disp=D, zero_disp=ZD, num_views=V, angle=18, **ALGO)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.d_dl = torch.empty((H, W), dtype=torch.float32, device=self.dev)
        self.d_dr = torch.empty_like(self.d_dl)
        self.d_out = torch.empty((H, W, 3), dtype=torch.uint8, device=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.h_dl = torch.empty((H, W), dtype=torch.float32).pin_memory()
        self.h_dr = torch.empty((H, W), dtype=torch.float32).pin_memory()
        self.h_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()

    def device_resident(self, frames, K, Wm):
        """frames/s with the inputs already in HBM: CUDA events around s2mv_process_sbs_device."""
        torch, pipe, sharding = self.torch, self.pipe, self.sharding
        d_frames = [torch.from_numpy(f).to(self.dev) for f in frames]
        NF = len(d_frames)

        def step(i):
            pipe.process_device(d_frames[i % NF].data_ptr(), 2 * W, self.d_dl.data_ptr(), self.d_dr.data_ptr(),
                                self.d_out.data_ptr(), self.stream)

        for i in range(Wm):
            step(i)
        torch.cuda.synchronize()
        pipe.enable_timing(True)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        stage_ms = {k: 0.0 for k in ("prepare", "costvol", "refine", "dibr")}
        kern_ms = {k: 0.0 for k in ("ci_h1", "v2", "v3", "h4_wta")}
        sharding.barrier()
        torch.cuda.synchronize()
        launches = 0
        for i in range(K):
            self.flush.fill_(i & 0xff)                 # L2 flush between timed iterations (not timed)
            ev[i][0].record()
            step(i)
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
        fps, _, slowest_s = sharding.aggregate_throughput(K, dev_ms / 1e3, self.dev)
        return {"fps": fps, "ms_per_step": 1e3 * slowest_s / K, "launches": launches,
                "stage_ms": {k: v / K for k, v in stage_ms.items()}, "kern_ms": {k: v / K for k, v in kern_ms.items()},
                "last_out": self.d_out.cpu().numpy()}

    def synchronous(self, np_ins, outs, K, Wm):
        """frames/s through s2mv_process_sbs (the adcensus_stm contract: returns after the D2H copies)."""
        pipe, sharding = self.pipe, self.sharding
        NF = len(np_ins)
        for i in range(Wm):
            pipe.adcensus_stm_into(np_ins[i % NF], *outs)
        sharding.barrier()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            pipe.adcensus_stm_into(np_ins[i % NF], *outs)
        self.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sharding.barrier()
        return sharding.aggregate_throughput(K, dt, self.dev)[0]

    def streamed(self, np_ins, K, Wm, depth=3, copy_into_slot=False):
        """frames/s through the asynchronous frame stream (the video loop): every frame page-locked host frame ->
        H2D -> all kernels -> D2H of both disparity maps and the interlaced frame into pinned host memory; copies of
        neighbouring frames overlap the kernels.  np_ins are page-locked: the library copies them to the device from
        where they lie.  copy_into_slot: the frame is first written into the slot's pinned buffer by the CPU (what a
        decoder that cannot write into page-locked memory of its own costs)."""
        pipe, sharding = self.pipe, self.sharding
        NF = len(np_ins)
        pipe.stream_open(depth)

        def run(n):
            got = None
            for i in range(n):
                if pipe.stream_pending == depth:
                    got = pipe.stream_collect(copy=False)
                if copy_into_slot:
                    np.copyto(pipe.stream_input_buffer(), np_ins[i % NF])   # stands in for a decoder writing the frame
                    pipe.stream_submit(None)
                else:
                    pipe.stream_submit(np_ins[i % NF])
            while pipe.stream_pending:
                got = pipe.stream_collect(copy=False)
            return got

        run(Wm)
        sharding.barrier()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = run(K)
        dt = time.perf_counter() - t0
        sharding.barrier()
        got = tuple(np.array(g) for g in got)
        pipe.stream_close()
        return sharding.aggregate_throughput(K, dt, self.dev)[0], got


def roofline_record(kern_ms, stage_ms):
    """Per-kernel roofline of the four cost-volume kernels + the record of the slowest one.

    `achieved` = the bytes the kernel's algorithm MUST move per launch (compulsory traffic of the fused pipeline:
    pass 1 only writes the volume, the fused vertical passes read and write it once, pass 4 only reads it:
    2V / 4V / 2V, V = one view's volume, both views per launch; 2V/4V/4V/2V when the vertical passes run as two
    launches) / its mean duration inside the timed steps.  ncu's DRAM bytes for the same
    kernels (profiles/ncu_traffic.json, `traffic`) agree with that figure to ~1 %, so `frac` is also the
    fraction of the measured HBM copy bandwidth the kernel really sustains.  `bound` comes from the ncu
    capture: "hbm" when the DRAM pipe is the busiest unit, "issue" when the instruction issue slots are.
    The stage-separable model of SURVEY §8(d) (40 B per disparity evaluation = 20V per frame: what the same
    passes would move unfused) is reported separately under `model_40B_per_de` and never called a fraction
    of anything the hardware did."""
    Vb = W * H * D * 4
    kern_ms = fold_fused(kern_ms)
    comp = {"ci_h1": 2 * Vb, "v2": 4 * Vb, "v3": 4 * Vb, "v2_v3_fused": 4 * Vb, "h4_wta": 2 * Vb}
    model = {"ci_h1": 6 * Vb, "v2": 4 * Vb, "v3": 4 * Vb, "v2_v3_fused": 8 * Vb, "h4_wta": 6 * Vb}
    peak, peak_src = measured_peak()
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    kernels = {}
    for k, ms in kern_ms.items():
        n = ncu.get(k, {})
        traffic = n.get("dram_bytes_per_launch")
        issue, dram, fma = n.get("issue_active_pct"), n.get("dram_throughput_pct"), n.get("fma_pipe_active_pct")
        bound = None if issue is None or dram is None else ("hbm" if dram >= max(issue, fma or 0.0) else "issue")
        kernels[k] = {"kernel": KERNEL_NAMES[k], "ms_per_launch": ms, "compulsory_bytes_per_launch": comp[k],
                      "achieved_gbs": comp[k] / (ms * 1e-3) / 1e9, "frac": comp[k] / (ms * 1e-3) / 1e9 / peak,
                      "traffic": traffic,
                      "actual_dram_gbs": None if traffic is None else traffic / (ms * 1e-3) / 1e9,
                      "actual_dram_frac": None if traffic is None else traffic / (ms * 1e-3) / 1e9 / peak,
                      "ncu_issue_active_pct": issue, "ncu_dram_throughput_pct": dram, "ncu_fma_pipe_active_pct": fma,
                      "bound": bound}
    dom = max(kern_ms, key=kern_ms.get)
    kd = kernels[dom]
    total_ms = sum(kern_ms.values())
    comp = {k: comp[k] for k in kern_ms}
    costvol_ms = stage_ms["prepare"] + stage_ms["costvol"]
    de = 2.0 * W * H * D
    return {"bound": kd["bound"] or "hbm", "kernel": kd["kernel"], "achieved": kd["achieved_gbs"], "peak": peak,
            "unit": "GB/s", "frac": kd["frac"], "traffic": kd["traffic"], "actual_dram_frac": kd["actual_dram_frac"],
            "peak_source": peak_src, "algorithmic_bytes_per_launch": comp[dom], "ms_per_launch": kd["ms_per_launch"],
            "kernels": kernels,
            "costvol_four_kernels": {"ms": total_ms, "compulsory_bytes": sum(comp.values()),
                                     "achieved_gbs": sum(comp.values()) / (total_ms * 1e-3) / 1e9,
                                     "frac": sum(comp.values()) / (total_ms * 1e-3) / 1e9 / peak},
            "model_40B_per_de": {"note": "stage-separable model (SURVEY 8d): 20V per frame as if no stage were fused; "
                                         "a throughput expressed in bytes, not traffic",
                                 "bytes_per_launch": model[dom], "costvol_leg_ms": costvol_ms,
                                 "mde_per_s": de / (costvol_ms * 1e-3) / 1e6,
                                 "equivalent_gbs": 40.0 * de / (costvol_ms * 1e-3) / 1e9,
                                 "equivalent_over_peak": 40.0 * de / (costvol_ms * 1e-3) / 1e9 / peak}}


def reference_gpu_record():
    """The reference's own CUDA kernels on this box (oracle/_ref, built from /root/reference by oracle/build_ref.sh),
    timed by oracle/ref_gpu_time.py in a separate process: a reported baseline, never on the product path."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_gpu_time.py")], capture_output=True,
                           text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": (r.stderr or r.stdout).strip().splitlines()[-1:] or ["no output"]}
        return json.loads(line[-1])
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}


def rowband_record(steps, warmup):
    """N > 1: BASELINE config 4 (one 3840x2160 D=256 frame in N row bands, halo rows over NVLink) -- the
    multi-GPU mode with communication on the data path -- next to the frame-parallel number."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import rowband_bench
    a = types.SimpleNamespace(height=2160, width=3840, disp=256, seed=4000, steps=steps, warmup=warmup, check=True,
                              baseline=True, transport="p2p", sha=False, single=False)
    try:
        res = rowband_bench.measure(a)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)}
    if res is None:
        return None
    chk = res.get("check", {})
    return {"workload": res["workload"], "ms_per_frame": res["ms_per_frame"],
            "single_context_ms": res.get("single_context_ms_per_frame"),
            "speedup_vs_single_context": res.get("speedup_vs_single_context"),
            "bit_exact": bool(chk) and all(chk.values()), "check": chk,
            "phase_ms": res["phase_ms_max_over_ranks"], "transport": res["transport"],
            "band_rows": res["band_rows"], "sub_image_rows": res["sub_image_rows"], "halo_rows": res["halo_rows"],
            "vertical_passes": res.get("vertical_passes"), "arena_gb_per_gpu": res["arena_gb_per_gpu"], "steps": res["steps"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-binding", action="store_true", help="N > 1: leave the ranks on the default CPU set")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.* (other workloads, reference GPU timing, row bands)")
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
    from s2mv_b200_pkg import sharding, synth

    rank, world, local_rank = sharding.dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    # several ranks on one host: each runs on the CPUs next to its GPU, so its pinned frame slots and the copies into
    # them stay on that socket (one rank keeps all cores: it also times the CPU baseline)
    cpus = sharding.bind_to_gpu_cpus(local_rank) if world > 1 and not args.no_cpu_binding else None
    sharding.init_process_group("nccl")
    K, Wm = args.steps, max(args.warmup, 3)

    frames, workload = load_frames(args.workload, rank)
    sbs = frames[0]
    rig = Rig(torch, s2mv_b200, sharding, local_rank)
    pipe = rig.pipe
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident throughput (the `value`) -------------------------
    dres = rig.device_resident(frames, K, Wm)

    # ---- end to end through the host-buffer C-ABI call --------------------
    h_frames = [torch.from_numpy(f).pin_memory() for f in frames]
    np_ins = [h.numpy() for h in h_frames]
    np_dl, np_dr, np_out = rig.h_dl.numpy(), rig.h_dr.numpy(), rig.h_out.numpy()
    e2e_sync_fps = rig.synchronous(np_ins, (np_dl, np_dr, np_out), K, Wm)
    ref_out = np.array(np_out)
    if len(frames) == 1:
        assert np.array_equal(ref_out, dres["last_out"]), "host and device entry points disagree"

    # ---- the same call with PAGEABLE caller buffers (what an unchanged video_io.cpp passes) ----
    pg_in = [np.array(f) for f in frames]
    pg = (np.empty((H, W), np.float32), np.empty((H, W), np.float32), np.empty((H, W, 3), np.uint8))
    pageable = {}
    for mode, m in (("default", 2), ("staged", 0), ("registered", 1)):   # default = what the adcensus_stm shim runs
        pipe.set_host_registration(m)
        pageable[mode] = rig.synchronous(pg_in, pg, K, Wm)
        assert np.array_equal(pg[2], ref_out)
    pipe.set_host_registration(0)

    # ---- end to end through the asynchronous frame stream (the video loop) ----
    DEPTH = 3
    e2e_fps, got = rig.streamed(np_ins, K, Wm, DEPTH)
    e2e_slot_fps, got2 = rig.streamed(np_ins, K, Wm, DEPTH, copy_into_slot=True)
    clocks = sampler.stop()
    assert np.array_equal(got2[2], ref_out)
    assert np.array_equal(got[2], ref_out) and np.array_equal(got[0], np_dl), "stream and synchronous entry points disagree"

    # ---- other workloads, same code, fewer steps (reported under `extra`, never the headline) ----
    extra = {}
    if not args.no_extras and world == 1:
        Ke = max(3, min(K, 10))
        others = {}
        if args.workload != "config3":
            others["config3"] = (load_frames("config3", rank)[0], WORKLOAD3)
        bud = np.load(os.path.join(ROOT, "tests", "golden", "bud_2_3.npz"))["sbs"]
        others["bud_1080p"] = ([np.ascontiguousarray(np.concatenate(
            [synth.upscale_bilinear(bud[:, :640], H, W), synth.upscale_bilinear(bud[:, 640:], H, W)], axis=1))],
            "img/bud_2+bud_3 bilinear-upscaled to 1920x1080 (a NON-degenerate bundled pair), D=128, zd=64, 8 views")
        for name, (fr, wl) in others.items():
            r = rig.device_resident(fr, Ke, 3)
            pin = [torch.from_numpy(f).pin_memory() for f in fr]
            e, _ = rig.streamed([h.numpy() for h in pin], Ke, 3, DEPTH)
            extra[name] = {"workload": wl, "value": r["fps"], "e2e": e, "unit": "frames/s", "steps": Ke,
                           "stage_ms": r["stage_ms"], "costvol_kernel_ms": fold_fused(r["kern_ms"])}
    pipe.close()
    if not args.no_extras and world > 1:
        rb = rowband_record(5, 2)
        if rb is not None:
            extra["rowband"] = rb

    if rank != 0:
        return 0

    roofline = roofline_record(dres["kern_ms"], dres["stage_ms"])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_frame(sbs)                                    # warm (page-in, thread pool)
        dt, cores = cpu_frame(sbs)
        cpu = {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE,
               "band_x8_estimate_frames_per_s": 1.0 / cpu_band_x8(sbs)}
    if not args.no_extras and world == 1:
        extra["reference_gpu"] = reference_gpu_record()

    print(json.dumps({
        "metric": METRIC, "value": dres["fps"], "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": dres["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": DATA[args.workload],
        "config": config_dict(workload, world),
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(sbs.nbytes),
                "d2h_bytes_per_step": int(np_dl.nbytes + np_dr.nbytes + np_out.nbytes),
                "api": f"s2mv_stream_submit/collect, {DEPTH} frames in flight (page-locked host frame -> H2D -> kernels -> D2H into pinned host memory)",
                "with_cpu_copy_into_pinned_slot": {"value": e2e_slot_fps, "unit": "frames/s",
                                                   "api": "the frame is first written into the slot's pinned buffer by the CPU (s2mv_stream_input_buffer), as a decoder would"},
                "synchronous_call": {"value": e2e_sync_fps, "unit": "frames/s", "api": "s2mv_process_sbs (adcensus_stm contract), pinned caller buffers",
                                     "pageable_caller_buffers": {"shim_default_auto_page_lock": pageable["default"],
                                                                 "staged_through_pinned": pageable["staged"],
                                                                 "page_locked_in_place": pageable["registered"],
                                                                 "unit": "frames/s"}}},
        "gpu_launches": dres["launches"],
        "host_cpus_per_rank": None if cpus is None else len(cpus),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reference_gpu": extra.get("reference_gpu"),
        "stage_ms": dres["stage_ms"],
        "costvol_kernel_ms": fold_fused(dres["kern_ms"]),
        "extra": extra,
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
