#!/usr/bin/env python
"""ref_gpu_time.py — time the REFERENCE's own CUDA kernels (built for sm_100 by oracle/build_ref.sh into
oracle/_ref/) on this box.  TEST / MEASUREMENT INFRASTRUCTURE, not the product: bench.py runs it as a
separate process for the `reference_gpu` record of its JSON line (SURVEY §8d "Reference GPU baseline").

  config 1 geometry (640x384, D=64, in the reference's validity domain): libs2mv_ref.so, unmodified;
  config 2 geometry (1920x1080, D=128): libs2mv_ref_patched.so — kernel bodies unchanged, launch geometry
  patched (build_ref.sh P1-P6, listed in `deviations`).

Each is timed with CUDA events around adcensus_stm (d_io.cu:7-238: H2D, all kernels, D2H, and its >30
cudaMalloc/cudaFree pairs per frame), warm, best of N — "as shipped" — and, for the patched build, again with
the allocations served from a size-keyed cache ("allocations hoisted").  Prints one JSON object.
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

DEVIATIONS = [
    "P1 d_ca_cross.cu:190 construction kernel block (num_cols,1)->(160,1)",
    "P2 d_ca_cross.cu:258,267 cost_transpose_kernel_4 -> the reference's bounds-checked cost_transpose_kernel, ceil grid",
    "P3 d_dc_wta.cu:45 block (num_cols,1)->(160,1)",
    "P4 d_dr_dcc.cu:94 block (num_cols,1)->(160,1)",
    "P5 d_dr_irv.cu:184-206 dhist[65]->dhist[512], max(num_disp,65) bins; :169 Q15 barrier; :248 block 32x32->32x24",
    "P6 d_filter_gaussian.cu:140 block 32x32->32x24",
    "T five D2H stage taps in d_io.cu (inactive while timing)",
    "A cudaMalloc/cudaFree routed through ref_harness.cu (pass-through for ms_as_shipped, cached for ms_alloc_hoisted)",
]


def p(a):
    return C.c_void_p(a.ctypes.data)


def f(x):
    return C.c_float(float(x))


def time_stm(lib, sbs, H, W, D, zd, warmup, iters):
    dl = np.zeros((H, W), np.float32)
    dr = np.zeros_like(dl)
    out = np.zeros((H, W, 3), np.uint8)
    lib.ref_time_adcensus_stm.restype = C.c_float
    return float(lib.ref_time_adcensus_stm(p(sbs), p(dl), p(dr), p(out), H, sbs.shape[1], W, H, W, 3, 8, 18, D, zd,
                                           f(10.0), f(30.0), f(20.0), f(6.0), 17, 9, 20, f(0.4), warmup, iters))


def main():
    import s2mv_b200  # noqa: F401  (package alias only; nothing of the product runs here)
    from s2mv_b200_pkg import synth
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    res = {}
    fish = np.load(os.path.join(ROOT, "tests", "golden", "fish_1_2.npz"))["sbs"]
    bud = np.load(os.path.join(ROOT, "tests", "golden", "bud_2_3.npz"))["sbs"]
    so = os.path.join(HERE, "_ref", "libs2mv_ref.so")
    if os.path.exists(so):
        lib = C.CDLL(so)
        res["config1_640x384_d64"] = {"ms_as_shipped": time_stm(lib, np.ascontiguousarray(bud), 384, 640, 64, 32, 1, iters),
                                      "build": "unmodified reference, sm_100", "frame": "bud_2+bud_3"}
    so = os.path.join(HERE, "_ref", "libs2mv_ref_patched.so")
    if os.path.exists(so):
        lib = C.CDLL(so)
        for name, src in (("config2_fish_1080p_d128", fish), ("bud_1080p_d128", bud)):
            L = synth.upscale_bilinear(src[:, :640], 1080, 1920)
            R = synth.upscale_bilinear(src[:, 640:], 1080, 1920)
            sbs = np.ascontiguousarray(np.concatenate([L, R], axis=1))
            shipped = time_stm(lib, sbs, 1080, 1920, 128, 64, 1, iters)
            lib.ref_pool_enable(1)
            hoisted = time_stm(lib, sbs, 1080, 1920, 128, 64, 1, iters)
            lib.ref_pool_enable(0)
            res[name] = {"ms_as_shipped": shipped, "ms_alloc_hoisted": hoisted,
                         "build": "reference kernels, launch geometry patched (deviations)"}
        res["deviations"] = DEVIATIONS
    if not res:
        res["unavailable"] = "oracle/_ref/*.so not built (oracle/build_ref.sh needs /root/reference)"
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
