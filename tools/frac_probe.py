#!/usr/bin/env python
"""Timing probe of the fractional-pel refinement on a 1080p +-64 frame: integer search, then all 593 partitions of all 480
CTUs refined from the winners on the device.  Prints kernel times (CUDA events inside the library) as JSON."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import hm  # noqa: E402
from synth import frame_jobs, luma_frames, pad_plane  # noqa: E402

W, H, R, M = 1920, 1080, 64, 80
f = luma_frames(W, H, 2)
cur, ref = pad_plane(f[1], M, M, np.uint8), pad_plane(f[0], M, M, np.uint8)
me = hm.MotionEstimator(0, R)
me.set_lambda_q16(460000)
pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
me.upload(pc, cur); me.upload(pr, ref)
jobs = frame_jobs(W, H, R)
if len(sys.argv) > 1:                      # first N jobs only (the band one of N GPUs would get)
    jobs = jobs[:int(sys.argv[1])]
out = {}
for had in (1, 0):
    ms = []
    for it in range(6):
        me.search_frame_async(pc, pr, jobs, R)
        me.refine_frame(pc, pr, len(jobs), None, bool(had), asynchronous=True)
        me.sync()
        ms.append(me.last_frac_ms())
    out["had" if had else "sad"] = {"kernel_ms": [round(v, 4) for v in ms], "pus": len(jobs) * 593,
                                    "pu_per_s": len(jobs) * 593 / (min(ms) * 1e-3), "search_ms": round(me.last_kernel_ms(), 4)}
res = me.refine_frame(pc, pr, len(jobs), None, True)
out["winner_histogram"] = {str(k): int(v) for k, v in zip(*np.unique((res["mvx"] & 3) * 4 + (res["mvy"] & 3), return_counts=True))}
print(json.dumps(out))
