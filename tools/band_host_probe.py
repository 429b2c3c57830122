"""Host time per call of one band's end-to-end step (1080p +-64 cut for 8 GPUs: 60 jobs), on one GPU: where does a launch-bound step
spend its host microseconds?  Two contexts alternate like the group's two frame slots."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from _pkg import hm
from synth import frame_jobs, luma_frames, pad_plane
W, H, R, M = 1920, 1080, 64, 80
f = luma_frames(W, H, 2)
n_cur = torch.from_numpy(pad_plane(f[1], M, M)).pin_memory().numpy()
n_ref = torch.from_numpy(pad_plane(f[0], M, M)).pin_memory().numpy()
jobs = np.ascontiguousarray(frame_jobs(W, H, R)[:60])
lib = hm.HmmeLib.get()
cr, rr = lib.band_extent(jobs, R)
ctxs = []
NCTX = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for s in range(NCTX):
    me = hm.MotionEstimator(0, R); me.set_lambda_q16(460000)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    outs = [torch.zeros((60, 593), dtype=torch.int32).pin_memory().numpy().view(t) for t in (np.int32, np.int32, np.uint32, np.uint32)]
    ctxs.append((me, pc, pr, outs))
T = {"sync": 0.0, "upload_ref": 0.0, "upload_cur": 0.0, "search": 0.0, "fetch": 0.0}
K = 400
def step(s, acc):
    me, pc, pr, outs = ctxs[s % NCTX]
    t0 = time.perf_counter(); me.sync()
    t1 = time.perf_counter(); me.upload_rect(pr, n_ref, rr, M, M)
    t2 = time.perf_counter(); me.upload_rect(pc, n_cur, cr, M, M)
    t3 = time.perf_counter(); me.search_frame_async(pc, pr, jobs, R)
    t4 = time.perf_counter(); me.fetch_results(60, outs, asynchronous=True)
    t5 = time.perf_counter()
    if acc:
        for k, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            T[k] += d
for s in range(20):
    step(s, False)
t0 = time.perf_counter()
for s in range(K):
    step(s, True)
for c in ctxs:
    c[0].sync()
tot = (time.perf_counter() - t0) / K
print("contexts", NCTX, "step %.1f us; host time per call (us):" % (tot * 1e6), {k: round(v / K * 1e6, 1) for k, v in T.items()}, "kernel ms", ctxs[0][0].last_kernel_ms())
