"""Profiling driver: N frames of the bench workload (config 2) through the device-resident C-ABI call.
Used under ncu on the GPU box:  ncu ... python tools/prof_frame.py 4   (see profiles/README.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import s2mv_b200  # noqa: E402

sbs = bench.load_frame()
pipe = s2mv_b200.Pipeline(0, num_rows=bench.H, num_cols=bench.W, num_disp=bench.D, zero_disp=bench.ZD,
                          num_views=8, angle=18, **bench.ALGO)
d_sbs = torch.from_numpy(sbs).cuda()
d_dl = torch.empty((bench.H, bench.W), dtype=torch.float32, device="cuda")
d_dr = torch.empty_like(d_dl)
d_out = torch.empty((bench.H, bench.W, 3), dtype=torch.uint8, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    pipe.process_device(d_sbs.data_ptr(), 2 * bench.W, d_dl.data_ptr(), d_dr.data_ptr(), d_out.data_ptr(), 0)
torch.cuda.synchronize()
print("frames", n, "launches/frame", pipe.last_launch_count)
