"""Small driver for compute-sanitizer: one packed-kernel frame batch (multi-tile, ragged last round) and one 16-bit call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _pkg import hm
from synth import frame_jobs, luma_frames, pad_plane
W, H, R = 128, 64, 20
f = luma_frames(W, H, 2, seed=3)
M = R + 8
cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
me = hm.MotionEstimator(0, 64)
me.set_lambda_q16(460000)
pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
me.upload(pc, cur); me.upload(pr, ref)
me.search_frame(pc, pr, frame_jobs(W, H, R), R)
me.search_ctu((2 * cur[M:M + 64, M:M + 64].astype(np.int32) - 7).astype(np.int16), ref, 0, 0, M, M, 4, -4, -4)
print("sanitize_smoke done")
