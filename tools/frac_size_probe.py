#!/usr/bin/env python
"""Where the fractional-pel refinement kernel spends its time: one PU size at a time, the 1080p frame tiled with PUs of that size at
random integer MVs.  Prints ms, ns per PU and ns per 8x8 tile-iteration for every size, plus a checksum of the results (so that
kernel variants can be compared without the oracle).  Usage: frac_size_probe.py [sad]"""
import json
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import hm  # noqa: E402
from synth import luma_frames, pad_plane  # noqa: E402

W, H, R, M = 1920, 1080, 64, 80
use_had = "sad" not in sys.argv[1:]
only = [a for a in sys.argv[1:] if "x" in a]          # e.g. 64x64: that size only (profiling runs)
f = luma_frames(W, H, 2)
cur, ref = pad_plane(f[1], M, M, np.uint8), pad_plane(f[0], M, M, np.uint8)
me = hm.MotionEstimator(0, R)
me.set_lambda_q16(460000)
pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
me.upload(pc, cur); me.upload(pr, ref)
rng = np.random.default_rng(7)
out = {}
for (w, h) in [(8, 8), (8, 4), (4, 8), (16, 8), (8, 16), (16, 4), (16, 12), (12, 16), (16, 16), (32, 8), (32, 32), (64, 16), (64, 64)]:
    if only and "%dx%d" % (w, h) not in only:
        continue
    xs, ys = np.meshgrid(np.arange(0, W - w + 1, w), np.arange(0, H - h + 1, h))
    n = xs.size
    reps = max(1, 150000 // n)                 # small grids are repeated so that every launch fills the GPU
    pus = np.zeros((n * reps, 8), np.int32)
    pus[:, 0], pus[:, 1] = np.tile(xs.ravel(), reps), np.tile(ys.ravel(), reps)
    pus[:, 2], pus[:, 3] = w, h
    pus[:, 4:6] = rng.integers(-32, 33, (n * reps, 2))
    ms = []
    for it in range(4):
        res = me.refine_frac(pc, pr, pus, use_had)
        ms.append(me.last_frac_ms())
    tiles = ((w + 7) // 8) * ((h + 7) // 8)
    t = min(ms)
    out["%dx%d" % (w, h)] = {"pus": int(n * reps), "ms": round(t, 4), "ns_per_pu": round(t * 1e6 / (n * reps), 2),
                             "ns_per_tile": round(t * 1e6 / (n * reps * tiles), 2), "crc": zlib.crc32(res.tobytes())}
print(json.dumps(out, indent=1))
