"""BASELINE config 5: cost-volume stress (ADCensus + cross aggregation H,V,V,H + WTA, both views) on a
synthetic frame, swept over num_disp, through the device-pointer C-ABI call s2mv_costvol_device.

  python tools/costvol_sweep.py [--height 4320 --width 7680 --disps 64,128,256,512 --steps 3 --check-rows 48]

One JSON line per num_disp: ms per frame (CUDA events, L2 flushed between steps), MDE/s
(2*W*H*D / t), the 40 B/DE algorithmic bandwidth against the measured HBM peak, arena bytes
and whether the volumes were chunk-sequential.  --check-rows R: the WTA disparities of R image rows are
compared bit for bit with the CPU oracle run on that band plus a (2*usd + 4)-row apron (the aggregation's
vertical reach plus the census window), with the GPU's own exponential tables.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import s2mv_b200  # noqa: E402
from s2mv_b200_pkg import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--seed", type=int, default=8000)
    ap.add_argument("--disps", default="64,128,256,512")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check-rows", type=int, default=0)
    args = ap.parse_args()
    H, W = args.height, args.width
    sbs = synth.make_sbs(H, W, args.seed)
    dev = torch.device("cuda", 0)
    d_sbs = torch.from_numpy(sbs).to(dev)
    d_dl = torch.empty((H, W), dtype=torch.float32, device=dev)
    d_dr = torch.empty_like(d_dl)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, peak_src = bench.measured_peak()
    st = torch.cuda.current_stream().cuda_stream
    usd = bench.ALGO["usd"]
    for D in [int(x) for x in args.disps.split(",")]:
        zd = D // 2
        with s2mv_b200.Pipeline(0, num_rows=H, num_cols=W, num_disp=D, zero_disp=zd, num_views=8, angle=18,
                                **bench.ALGO) as pipe:
            for _ in range(args.warmup):
                pipe.costvol_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), st)
            torch.cuda.synchronize()
            ms = 0.0
            for i in range(args.steps):
                flush.fill_(i)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                pipe.costvol_device(d_sbs.data_ptr(), 2 * W, d_dl.data_ptr(), d_dr.data_ptr(), st)
                b.record()
                b.synchronize()
                ms += a.elapsed_time(b)
            ms /= args.steps
            de = 2.0 * W * H * D
            row = {"workload": f"config5: synthetic {W}x{H} seed {args.seed}, cost volume only (prepare + CI + 4 passes + WTA)",
                   "num_disp": D, "zero_disp": zd, "ms_per_frame": ms, "mde_per_s": de / (ms * 1e-3) / 1e6,
                   "algorithmic_gbs_40B_per_de": 40.0 * de / (ms * 1e-3) / 1e9, "peak_gbs": peak, "peak_source": peak_src,
                   "frac": 40.0 * de / (ms * 1e-3) / 1e9 / peak, "arena_gb": pipe.arena_bytes / 1e9,
                   "chunk_sequential": pipe.chunk_sequential, "launches": pipe.last_launch_count, "steps": args.steps}
            if args.check_rows:
                import oracle_py
                R, ap_rows = args.check_rows, 2 * usd + 4  # aggregation reaches 2*usd rows, the census window 3 more
                y0 = (H - R) // 2
                band = np.ascontiguousarray(sbs[y0 - ap_rows:y0 + R + ap_rows])
                odl, odr = oracle_py.costvol(np.ascontiguousarray(band[:, :W]), np.ascontiguousarray(band[:, W:2 * W]),
                                             D, zd, luts=pipe.exp_tables(), **{k: bench.ALGO[k] for k in
                                                                               ("ad_coeff", "census_coeff", "ucd", "lcd", "usd", "lsd")})
                gl, gr = d_dl[y0:y0 + R].cpu().numpy(), d_dr[y0:y0 + R].cpu().numpy()
                row["oracle_check"] = {"rows": [y0, y0 + R],
                                       "wta_left_equal": bool(np.array_equal(gl, odl[ap_rows:ap_rows + R])),
                                       "wta_right_equal": bool(np.array_equal(gr, odr[ap_rows:ap_rows + R]))}
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
