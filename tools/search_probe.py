#!/usr/bin/env python
"""Timing + parity probe of the integer search on the 1080p +-64 frame: kernel time (CUDA events inside the library), frame time with two
contexts alternating, and a full comparison with the oracle.  HMME_B200_LIB selects the build of the library under test.
    python tools/search_probe.py [--no-check] [--range R]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkg import hm  # noqa: E402
from synth import frame_jobs, luma_frames, pad_plane  # noqa: E402

R = int(sys.argv[sys.argv.index("--range") + 1]) if "--range" in sys.argv else 64
W, H, M = 1920, 1080, 80 if R <= 64 else R + 16
f = luma_frames(W, H, 2)
cur, ref = pad_plane(f[1], M, M, np.uint8), pad_plane(f[0], M, M, np.uint8)
jobs = frame_jobs(W, H, R)
ctx = [hm.MotionEstimator(0, R) for _ in range(2)]
planes = []
for me in ctx:
    me.set_lambda_q16(460000)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    planes.append((pc, pr))
ms = []
for it in range(8):
    ctx[0].search_frame_async(planes[0][0], planes[0][1], jobs, R)
    ctx[0].sync()
    ms.append(ctx[0].last_kernel_ms())
import torch  # noqa: E402
K = 20
for me in ctx:
    me.sync()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(K):
    ctx[it & 1].search_frame_async(planes[it & 1][0], planes[it & 1][1], jobs, R)
for me in ctx:
    me.sync()
t1 = time.perf_counter()
out = {"lib": os.environ.get("HMME_B200_LIB", "default"), "range": R, "kernel_ms_min": round(min(ms), 4), "kernel_ms_med": round(float(np.median(ms)), 4),
       "two_ctx_ms_per_frame_wall": round((t1 - t0) * 1e3 / K, 4)}
if "--no-check" not in sys.argv:
    from oracle.pyoracle import Oracle
    got = ctx[0].search_frame(planes[0][0], planes[0][1], jobs, R)
    want = Oracle().search_frame(cur.astype(np.int16), (M, M), ref.astype(np.int16), (M, M), jobs, R, 460000, nthreads=16)
    out["mismatches"] = int(sum(int(np.count_nonzero(g != w)) for g, w in zip(got, want)))
print(json.dumps(out))
