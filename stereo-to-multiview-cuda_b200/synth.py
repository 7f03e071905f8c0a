"""Synthetic stereo frames (BASELINE configs 3-5): piecewise-smooth
content so that cross arms have realistic lengths (white noise would give arm
length 1 everywhere and make aggregation trivially cheap).

left  = sum of low-frequency sinusoid gradients per channel + random filled
        ellipses + Gaussian noise (sigma 2), clipped to u8
right = left warped horizontally by a smooth disparity field in [-24, 40] px
        plus a per-ellipse offset, re-noised
Deterministic in (height, width, seed): numpy.random.default_rng(seed).
"""
import numpy as np


def make_pair(height, width, seed, n_ellipses=40, noise_sigma=2.0, disp_lo=-24.0, disp_hi=40.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    sy, sx = yy / height, xx / width
    img = np.zeros((height, width, 3), np.float32)
    for c in range(3):
        acc = np.full((height, width), 110.0, np.float32)
        for _ in range(8):
            fx, fy = rng.uniform(0.3, 3.0, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += rng.uniform(8, 28) * np.sin(2 * np.pi * (fx * sx + fy * sy) + ph).astype(np.float32)
        img[..., c] = acc
    # smooth background disparity
    disp = np.zeros((height, width), np.float32)
    for _ in range(3):
        fx, fy = rng.uniform(0.2, 1.2, 2)
        ph = rng.uniform(0, 2 * np.pi)
        disp += np.sin(2 * np.pi * (fx * sx + fy * sy) + ph).astype(np.float32)
    disp = (disp - disp.min()) / max(float(disp.max() - disp.min()), 1e-6)
    disp = disp_lo + 0.5 * (disp_hi - disp_lo) * disp
    scale = min(height, width)
    for _ in range(n_ellipses):
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        ax, ay = rng.uniform(0.03, 0.18, 2) * scale
        th = rng.uniform(0, np.pi)
        # same ellipse test as on the full grid, evaluated inside its bounding box only
        r = float(max(ax, ay)) + 1.0
        x0, x1 = max(int(cx - r), 0), min(int(cx + r) + 2, width)
        y0, y1 = max(int(cy - r), 0), min(int(cy + r) + 2, height)
        colour = rng.uniform(20, 235, 3).astype(np.float32)
        dval = rng.uniform(0.5 * (disp_lo + disp_hi), disp_hi)
        if x0 >= x1 or y0 >= y1:
            continue
        dx, dy = xx[y0:y1, x0:x1] - cx, yy[y0:y1, x0:x1] - cy
        u = dx * np.cos(th) + dy * np.sin(th)
        v = -dx * np.sin(th) + dy * np.cos(th)
        m = (u / ax) ** 2 + (v / ay) ** 2 <= 1.0
        img[y0:y1, x0:x1][m] = colour
        disp[y0:y1, x0:x1][m] = dval
    left = img + rng.normal(0, noise_sigma, img.shape).astype(np.float32)
    # right(x) = left(x + d): sample the clean image at shifted columns (nearest)
    src = np.clip(np.rint(xx + disp), 0, width - 1).astype(np.int64)
    right = np.take_along_axis(img, src[..., None].repeat(3, axis=2), axis=1)
    right = right + rng.normal(0, noise_sigma, img.shape).astype(np.float32)
    to_u8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)  # noqa: E731
    return to_u8(left), to_u8(right)


def make_sbs(height, width, seed, **kw):
    left, right = make_pair(height, width, seed, **kw)
    return np.ascontiguousarray(np.concatenate([left, right], axis=1))


def upscale_bilinear(img, out_h, out_w):
    """Bilinear resize following the reference's own formula (tx_scale_bilinear_kernel,
    d_tx_scale.cu:30-52: sample at (tx/W_out)*W_in, floor + fractional weights, clamped)."""
    in_h, in_w = img.shape[:2]
    xs = np.minimum((np.arange(out_w, dtype=np.float32) / np.float32(out_w)) * np.float32(in_w), in_w - 1)
    ys = np.minimum((np.arange(out_h, dtype=np.float32) / np.float32(out_h)) * np.float32(in_h), in_h - 1)
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    x1 = np.minimum(x0 + 1, in_w - 1)
    y1 = np.minimum(y0 + 1, in_h - 1)
    wx = (xs - x0).astype(np.float32)[None, :, None]
    wy = (ys - y0).astype(np.float32)[:, None, None]
    f = img.astype(np.float32)
    top = f[y0][:, x0] * (1 - wx) + f[y0][:, x1] * wx
    bot = f[y1][:, x0] * (1 - wx) + f[y1][:, x1] * wx
    return (top * (1 - wy) + bot * wy).astype(np.uint8)
