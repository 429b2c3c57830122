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
pp = [P(), P(), P()]
u8_cur = torch.from_numpy(pad_plane(f[1], M, M, np.uint8)).pin_memory().numpy()
u8_ref = torch.from_numpy(pad_plane(f[0], M, M, np.uint8)).pin_memory().numpy()
def run(name, up, search, fetch, K=60, two=True, nctx=2):
    for p in pp: p.me.sync()
    t0 = time.perf_counter()
    for s in range(K):
        p = pp[s % nctx] if two else pp[0]
        p.me.sync()
        if up == 1:
            p.me.upload(p.pr, n_ref, asynchronous=True); p.me.upload(p.pc, n_cur, asynchronous=True)
        elif up == 2:
            p.me.upload(p.pr, u8_ref); p.me.upload(p.pc, u8_cur)
        elif up == 3:
            p.me.upload(p.pr, n_ref, asynchronous=True)
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
run("all, 3 ctx", 1, 1, 1, nctx=3)
run("all, 2 ctx again", 1, 1, 1)
run("search only, 2 ctx again", 0, 1, 0)
run("u8 sync upload+search, 2 ctx", 2, 1, 0)
run("one s16 upload+search, 2 ctx", 3, 1, 0)
run("upload+search, 2 ctx again", 1, 1, 0)
# --- does a concurrent, unrelated H2D stream slow the search kernel?
side = torch.cuda.Stream()
hbuf = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
def run_bg(name, K=60):
    for p in pp: p.me.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(K):
        p = pp[s & 1]
        p.me.sync()
        with torch.cuda.stream(side):
            dbuf.copy_(hbuf, non_blocking=True)         # 16 MB, independent of the search
        p.me.search_frame_async(p.pc, p.pr, jobs, R)
    for p in pp: p.me.sync()
    torch.cuda.synchronize()
    print("%-34s %.3f ms/step" % (name, (time.perf_counter() - t0) * 1e3 / K))
run_bg("search + unrelated 16MB H2D")
run("search only, 2 ctx (ref)", 0, 1, 0)
# --- kernel duration (CUDA events around the search kernel) with / without concurrent H2D traffic
def kern_ms(bg, K=30):
    p = pp[0]; ms = []
    for s in range(K):
        p.me.sync(); torch.cuda.synchronize()
        if bg:
            with torch.cuda.stream(side):
                for _ in range(4): dbuf.copy_(hbuf, non_blocking=True)      # 64 MB in flight = ~1.2 ms of PCIe
        p.me.search_frame_async(p.pc, p.pr, jobs, R)
        ms.append(p.me.last_kernel_ms())
    torch.cuda.synchronize()
    return sum(ms[5:]) / len(ms[5:])
print("search kernel ms: alone %.3f, with concurrent H2D %.3f" % (kern_ms(False), kern_ms(True)))
