"""Regenerates the committed input fixtures from the reference's bundled images
(run in the build container, where /root/reference exists; the GPU box only
ever reads the committed .npz files).

  bud_2_3.npz   img/bud_2.bmp (left) + img/bud_3.bmp (right), 640x384  -> BASELINE config 1 (in-domain stand-in)
  fish_1_2.npz  img/fish_1.bmp + img/fish_2.bmp, 640x384                -> BASELINE config 2 before upscaling
Each holds `sbs`: the side-by-side BGR frame (H, 2W, 3) uint8 exactly as cv2.imread decodes the BMPs
(OpenCV's imread is what image_io.cpp:95-96 uses).
"""
import os
import sys

import cv2
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
for name, (l, r) in {"bud_2_3": ("bud_2", "bud_3"), "fish_1_2": ("fish_1", "fish_2")}.items():
    L = cv2.imread(os.path.join(REF, "img", l + ".bmp"), cv2.IMREAD_COLOR)
    R = cv2.imread(os.path.join(REF, "img", r + ".bmp"), cv2.IMREAD_COLOR)
    assert L.shape == R.shape == (384, 640, 3)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), sbs=np.concatenate([L, R], axis=1))
    print(name, "ok")
