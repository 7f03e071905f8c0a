"""BASELINE config 4: ONE synthetic frame split into row bands over the GPUs of the box, one process
per GPU (torchrun), halos over NCCL point-to-point (NVLink), timed as single-frame latency.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/rowband_bench.py [--height 2160 --width 3840 --disp 256 --steps 5 --check]
  python tools/rowband_bench.py ...            (N = 1: the band is the whole frame, no exchange)

Prints one JSON line from rank 0: ms per frame (CUDA events on every rank, max over ranks, frame already
resident on every GPU), the single-context latency measured on rank 0 in the same run when the frame fits
one GPU (--baseline), and with --check the bit-for-bit comparison of the assembled outputs with that
single-context frame.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import s2mv_b200  # noqa: E402
from s2mv_b200_pkg import rowband, sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--disp", type=int, default=256)
    ap.add_argument("--seed", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="compare with the single-context frame on rank 0")
    ap.add_argument("--baseline", action="store_true", help="also time the single-context frame on rank 0")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="halo rows: peer stores from the producing kernels over CUDA IPC mappings, or NCCL send/recv")
    ap.add_argument("--vertical", default="auto", choices=["auto", "fused", "separate"],
                    help="the two vertical passes: one launch per band with one exchange of 2*usd rows (needs bands of 2*usd "
                         "rows), two launches with two exchanges of usd rows, or the library's choice (fused for tall bands)")
    ap.add_argument("--sha", action="store_true", help="print SHA-256 of the assembled outputs (to compare runs whose "
                                                       "single-context frame does not fit next to a band)")
    ap.add_argument("--single", action="store_true", help="N = 1 only: time the ordinary single-context frame "
                                                          "(chunk-sequential volumes when needed) instead of one band")
    args = ap.parse_args()
    res = measure(args)
    if res is not None:
        print(json.dumps(res), flush=True)
    return 0


def measure(args):
    """One row-band measurement; every rank calls it, rank 0 gets the result dict (others None).
    `args` needs: height width disp seed steps warmup check baseline transport sha single."""
    rank, world, local_rank = sharding.dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sharding.init_process_group("nccl")
    H, W, D = args.height, args.width, args.disp
    params = dict(num_rows=H, num_cols=W, num_disp=D, zero_disp=D // 2, num_views=8, angle=18, **bench.ALGO)
    sbs = synth.make_sbs(H, W, args.seed)
    d_sbs = torch.from_numpy(sbs).to(dev)

    if args.single:
        assert world == 1, "--single is the one-GPU reference run"
        import hashlib
        with s2mv_b200.Pipeline(local_rank, **params) as p:
            d_dl = torch.empty((H, W), dtype=torch.float32, device=dev)
            d_dr = torch.empty_like(d_dl)
            d_out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(args.warmup):
                p.process_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), st)
            torch.cuda.synchronize()
            t1 = 0.0
            for _ in range(args.steps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                p.process_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), st)
                b.record()
                b.synchronize()
                t1 += a.elapsed_time(b)
            res = {"workload": f"synthetic {W}x{H} seed {args.seed}, D={D}, ONE context (no bands), full pipeline",
                   "n_gpus": 1, "ms_per_frame": t1 / args.steps, "steps": args.steps, "arena_gb_per_gpu": p.arena_bytes / 1e9,
                   "chunk_sequential": p.chunk_sequential,
                   "sha": {k: hashlib.sha256(t.cpu().numpy().tobytes()).hexdigest() for k, t in
                           (("disp_l", d_dl), ("disp_r", d_dr), ("interlaced", d_out))}}
        return res

    band = rowband.DistBand(local_rank, rank, world, transport=args.transport,
                            fuse_vertical={"auto": None, "fused": True, "separate": False}[getattr(args, "vertical", "auto")],
                            **params)
    for _ in range(args.warmup):
        band.process(d_sbs, 2 * W)
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(args.steps):
        sharding.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out_l, out_r, out_i = band.process(d_sbs, 2 * W, check=False)
        b.record()
        b.synchronize()
        ms += sharding.reduce_scalar(a.elapsed_time(b), "max", dev)
    ms /= args.steps
    if band.transport == "p2p":
        band.check()                            # no wait ran out before its neighbour arrived
    phases = {}
    band.process(d_sbs, 2 * W, phases)          # one more frame with per-phase device times (this rank's)
    allph = [None] * world
    if world > 1:
        dist.all_gather_object(allph, phases)
    else:
        allph = [phases]
    arena = band.ctx.pipe.arena_bytes
    res = {"workload": f"config4-style: synthetic {W}x{H} seed {args.seed}, D={D}, one frame in {world} row band(s), "
                       "full pipeline (disparities + 8-view interlaced frame)",
           "n_gpus": world, "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "steps": args.steps,
           "band_rows": [y1 - y0 for y0, y1 in band.bands], "sub_image_rows": band.ctx.local_rows,
           "halo_rows": band.ctx.halo_rows, "vertical_passes": "fused" if band.ctx.halo_rows > params["usd"] else "separate", "halo_bytes_per_exchange_per_neighbour": 2 * band.ctx.halo_rows * W * 4 * (
               (D + 127) // 128 * 128 if D > 128 else 1 << (max(D, 4) - 1).bit_length()),
           "phase_ms_max_over_ranks": {k: max(p[k] for p in allph) for k in phases},
           "phase_ms_rank0": phases,
           "arena_gb_per_gpu": arena / 1e9, "transport": {"p2p": "halo rows stored by the producing kernels into CUDA-IPC-mapped neighbour volumes (NVLink), epoch words "
                                "on the stream; NCCL send/recv of the disparity rows each sub-image takes from other bands",
                         "nccl": "torch.distributed NCCL send/recv of the halo rows and of the disparity rows each sub-image takes from other bands",
                         "none": "one band"}[band.transport]}

    if args.sha:
        import hashlib
        sha = {}
        for k, t in (("disp_l", out_l), ("disp_r", out_r), ("interlaced", out_i)):
            full_t = rowband.allgather_rows_dist(t, band.bands, dist, torch) if world > 1 else t
            if rank == 0:
                sha[k] = hashlib.sha256(full_t.cpu().numpy().tobytes()).hexdigest()
            del full_t
        res["sha"] = sha
    if args.check or args.baseline:
        # assemble the frame on rank 0
        full = []
        for t in (out_l, out_r, out_i):
            if world > 1:
                full.append(rowband.allgather_rows_dist(t, band.bands, dist, torch))
            else:
                full.append(t)
        band.close()
        if rank == 0:
            with s2mv_b200.Pipeline(local_rank, **params) as p:
                d_dl = torch.empty((H, W), dtype=torch.float32, device=dev)
                d_dr = torch.empty_like(d_dl)
                d_out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
                st = torch.cuda.current_stream().cuda_stream
                for _ in range(args.warmup):
                    p.process_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), st)
                torch.cuda.synchronize()
                t1 = 0.0
                for _ in range(args.steps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    p.process_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), st)
                    b.record()
                    b.synchronize()
                    t1 += a.elapsed_time(b)
                res["single_context_ms_per_frame"] = t1 / args.steps
                res["speedup_vs_single_context"] = (t1 / args.steps) / ms
                if args.check:
                    res["check"] = {"disp_l_equal": bool(torch.equal(full[0], d_dl)),
                                    "disp_r_equal": bool(torch.equal(full[1], d_dr)),
                                    "interlaced_equal": bool(torch.equal(full[2], d_out))}
    else:
        band.close()
    sharding.barrier()
    return res if rank == 0 else None


if __name__ == "__main__":
    sys.exit(main())
