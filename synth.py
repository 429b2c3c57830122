"""Synthetic 8-bit luma frames shared by tests and bench.py (BASELINE.md section 4 generator: seeded uniform
canvas, 5x5 box blur, contrast x3 about 128, global pan (3,2) px/frame, one 48x48 inverted moving square),
plus HM-style padded planes (edge replication, TComPicYuv::extendPicBorder) and job lists."""
import numpy as np


def luma_frames(W, H, F, seed=1234):
    rng = np.random.default_rng(seed)
    big = rng.integers(0, 256, size=(H + 4 * F + 64, W + 4 * F + 64), dtype=np.uint8).astype(np.float32)
    k = 5
    c = np.cumsum(np.cumsum(np.pad(big, ((k, 0), (k, 0))), 0), 1)
    sm = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
    sm = np.clip((sm - 128) * 3 + 128, 0, 255).astype(np.uint8)
    out = []
    for t in range(F):
        y = sm[2 * t:2 * t + H, 3 * t:3 * t + W].copy()
        x0 = (40 + 7 * t) % (W - 48)
        y0 = (30 + 5 * t) % (H - 48)
        y[y0:y0 + 48, x0:x0 + 48] = 255 - y[y0:y0 + 48, x0:x0 + 48]
        out.append(y)
    return out


def pad_plane(luma, margin_x, margin_y, dtype=np.int16):
    """Edge-replicated padded plane; picture sample (0,0) at [margin_y, margin_x]."""
    return np.ascontiguousarray(np.pad(luma, ((margin_y, margin_y), (margin_x, margin_x)), mode="edge").astype(dtype))


def frame_jobs(W, H, R, pred=(0, 0), rows=None):
    """One job per FULL 64x64 CTU (partial boundary CTUs never run depth 0, TEncCu.cpp:424-425), raster order,
    window centred on `pred` (integer pel): lt = pred - R.  rows = (r0, r1) restricts to a CTU-row band."""
    nx, ny = W // 64, H // 64
    r0, r1 = rows if rows is not None else (0, ny)
    return np.array([[cx * 64, cy * 64, pred[0] - R, pred[1] - R] for cy in range(r0, r1) for cx in range(nx)], np.int32).reshape(-1, 4)
