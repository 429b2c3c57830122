"""Where does the pipelined end-to-end step lose time?  Times two-context alternation with subsets of the legs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from _pkg import hm
from synth import frame_jobs, luma_frames, pad_plane
W, H, R, M = 1920, 1080, 64, 80
f = luma_frames(W, H, 2)
n_cur = torch.from_numpy(pad_plane(f[1], M, M)).pin_memory().numpy()
n_ref = torch.from_numpy(pad_plane(f[0], M, M)).pin_memory().numpy()
jobs = frame_jobs(W, H, R)
class P:
    def __init__(s):
        s.me = hm.MotionEstimator(0, R); s.me.set_lambda_q16(460000)
        s.pc = s.me.alloc_plane(1, W, H, M, M); s.pr = s.me.alloc_plane(1, W, H, M, M)
        s.me.upload(s.pc, n_cur); s.me.upload(s.pr, n_ref)
        s.outs = [torch.zeros((len(jobs), 593), dtype=torch.int32).pin_memory().numpy().view(t) for t in (np.int32, np.int32, np.uint32, np.uint32)]
pp = [P(), P()]
def run(name, up, search, fetch, K=60, two=True):
    for p in pp: p.me.sync()
    t0 = time.perf_counter()
    for s in range(K):
        p = pp[s & 1] if two else pp[0]
        p.me.sync()
        if up:
            p.me.upload(p.pr, n_ref, asynchronous=True); p.me.upload(p.pc, n_cur, asynchronous=True)
        if search: p.me.search_frame_async(p.pc, p.pr, jobs, R)
        if fetch: p.me.fetch_results(len(jobs), p.outs, asynchronous=True)
    for p in pp: p.me.sync()
    print("%-34s %.3f ms/step" % (name, (time.perf_counter() - t0) * 1e3 / K))
run("search only, 2 ctx", 0, 1, 0)
run("search only, 1 ctx", 0, 1, 0, two=False)
run("upload only, 2 ctx", 1, 0, 0)
run("fetch only, 2 ctx", 0, 0, 1)
run("upload+search, 2 ctx", 1, 1, 0)
run("search+fetch, 2 ctx", 0, 1, 1)
run("all, 2 ctx", 1, 1, 1)
run("all, 1 ctx", 1, 1, 1, two=False)
