#!/usr/bin/env python
"""Timing / profiling probe of hmme_mc_cost: all 593 partitions of the 480 CTUs of a 1080p frame at random quarter-pel MVs."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import hm  # noqa: E402
from synth import frame_jobs, luma_frames, pad_plane  # noqa: E402

W, H, R, M = 1920, 1080, 64, 80
f = luma_frames(W, H, 2)
me = hm.MotionEstimator(0, R)
pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
me.upload(pc, pad_plane(f[1], M, M, np.uint8)); me.upload(pr, pad_plane(f[0], M, M, np.uint8))
jobs = frame_jobs(W, H, R)
rects = me.lib.partition_table()
rng = np.random.default_rng(7)
pus = np.zeros((len(jobs), 593, 6), np.int32)
pus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
pus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
pus[:, :, 2], pus[:, :, 3] = rects[None, :, 2], rects[None, :, 3]
pus[:, :, 4:6] = rng.integers(-4 * (R - 8), 4 * (R - 8), size=(len(jobs), 593, 2))
out = {}
for name, had in (("sad", False), ("hadamard", True)):
    t = []
    for _ in range(4):
        me.mc_cost(pc, pr, pus.reshape(-1, 6), had)
        t.append(round(me.last_frac_ms(), 4))
    out[name] = t
print(json.dumps(out))
