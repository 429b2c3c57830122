#!/usr/bin/env python
"""bench.py -- whole-CTU integer-pel motion estimation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p64|4k128|1080p64_ra|...] [--impl reference] [--no-verify]

A step = the integer ME of ONE frame against ONE reference picture: every full 64x64 CTU of the frame,
(2R+1)^2 candidates each, 593 partitions per candidate (config[1] of BASELINE.json at N=1: 1920x1080,
+-64 -> 480 jobs x 16641 candidates).  With N > 1 (torchrun, one rank per GPU) the frame's jobs are cut into
contiguous CTU-row bands by the LIBRARY (hmme_group_*, include/hmme_b200.h): every rank passes the same whole-frame
arguments to hmme_group_search_frame_async and gets its band's rows of the result tables back; the reference picture
reaches the GPUs either as band + halo rectangles over each GPU's own PCIe link (default) or by NCCL broadcast over
NVLink from rank 0 (`e2e.broadcast`, also measured).  Python only loops over steps.

Output: ONE JSON line on rank 0 (contract in the task statement): `value` = block-SAD evaluations/s with inputs
resident in HBM (K frames alternating over two of the group's frame slots = two streams, CUDA events around the whole
region, inputs cycled through more plane copies than fit in L2), `e2e` = the same metric through the group call with
pinned HOST planes (HM's int16 Pel; H2D of both planes' band rectangles + jobs, D2H of the four result arrays inside
the timed region), `roofline` = algorithmic integer lane-ops/s of the dominant kernel against the issue rate of both
integer pipes, 2 x the ALU rate measured live (`frac` = `frac_issue`; `frac_one_pipe` = the survey's one-pipe denominator), `verified` = every
CTU of every rank's band compared bit for bit with the CPU oracle after the timed regions, `per_ctu` = the
synchronous per-CTU call the encoder makes, `cpu_baseline` = the reference's own CPU full-search ME.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from synth import frame_jobs, luma_frames, pad_plane  # noqa: E402

WORKLOADS = {                      # name: (W, H, R)   -- BASELINE.json configs
    "1080p64": (1920, 1080, 64),   # config[1] (and the metric's "1080p frames/s (+-64)")
    "1080p64_ra": (1920, 1080, 64),  # config[2]: random access, two reference lists + bi-prediction refinement (+-4, 16-bit block)
    "4k128": (3840, 2160, 128),    # config[3]
    "1080p32": (1920, 1080, 32),
    "1080p16": (1920, 1080, 16),
    "416x240_64": (416, 240, 64),  # config[0] geometry
}
NPARTS = 593
INT_OPS_PER_CAND = 2803            # 1024 packed SADs + 593 adds + 593 cost adds + 593 min
INT_OPS_PER_CAND_16 = 3827         # same with 2048 packed 2x16-bit absolute differences (bi-prediction block)
PX_PER_CAND = 4096
LAMBDA_Q16 = 460000                # QP ~32 (SURVEY.md section 8d synthetic inputs)
BI_RANGE = 4                       # bipredSearchRange of the reference's configurations (cfg/encoder_randomaccess_main.cfg)
REF_ARM_BUDGET_S = 150            # wall-clock budget of one --impl reference run (see run_reference)
REF_ARM_CLIP = (416, 240)          # the reference arm's bounded sample: this crop of the workload's own frame pair (18 full CTUs, ~10 s on 16 cores)


def workload_geometry(name):
    W, H, R = WORKLOADS[name]
    margin = max(80, R + 16)       # HM pads by 80 (TComPicYuv.cpp:93-94); larger ranges need more for in-bounds windows
    return W, H, R, margin


def shared_config(name):
    """`config` of the JSON line: identical in the b200 arm and the reference arm (the driver compares them)."""
    W, H, R = WORKLOADS[name]
    ncx, ncy = W // 64, H // 64
    return {"workload": "%s: %dx%d luma, 64x64 CTU, integer-pel full search +-%d, 1 reference picture, %d CTU jobs x %d candidates x 593 partitions"
                        % (name, W, H, R, ncx * ncy, (2 * R + 1) ** 2),
            "lambda_q16": LAMBDA_Q16,
            "reference_arm_sample": "the reference arm (--impl reference) times the reference's CPU full search (+-%d) on the top-left %dx%d crop of this "
                                    "workload's frame pair per step, one single-threaded encoder process per host core; rates are per block-SAD evaluation"
                                    % (R, REF_ARM_CLIP[0], REF_ARM_CLIP[1])}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one long-running
    `nvidia-smi -lms` process, started before the region and stopped after it; samples outside [t0, t1] are dropped."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.t = [], []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        if not self.proc:
            return
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 8:
                self.rows.append(c[1:])
                self.t.append(time.perf_counter())

    def start(self):
        t_end = time.perf_counter() + 3.0      # let the first samples arrive before the region starts (nvidia-smi takes a while to come up)
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.02)
        self.t0 = time.perf_counter()

    def summary(self):
        t1 = time.perf_counter()
        time.sleep(0.05)
        n0 = len(self.rows)
        t_end = time.perf_counter() + 1.0      # a region shorter than the sampling period: take the sample that follows it
        while self.proc and len(self.rows) == n0 and not any(self.t0 <= t for t in self.t) and time.perf_counter() < t_end:
            time.sleep(0.02)
        if self.proc:
            self.proc.terminate()
        self.thread.join(timeout=3)
        rows = [r for r, t in zip(self.rows, self.t) if self.t0 <= t <= t1 + 0.03] or self.rows[-3:]   # none inside: the nearest ones
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "samples": len(sm),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons}


def cpu_oracle_throughput(W, H, R, margin, njobs_sample, threads):
    """Times the CPU oracle port (oracle/hmme_oracle.c, hierarchical variant) on a bounded sample of the workload's jobs."""
    from oracle.pyoracle import Oracle
    f = luma_frames(W, H, 2)
    cur, ref = pad_plane(f[1], margin, margin), pad_plane(f[0], margin, margin)
    jobs = frame_jobs(W, H, R)
    pick = np.linspace(0, len(jobs) - 1, num=min(njobs_sample, len(jobs))).astype(int)
    o = Oracle()
    t0 = time.perf_counter()
    o.search_frame(cur, (margin, margin), ref, (margin, margin), jobs[pick], R, LAMBDA_Q16, nthreads=threads)
    dt = time.perf_counter() - t0
    cands = len(pick) * (2 * R + 1) ** 2
    return cands * NPARTS / dt, dt, len(pick)


FRAC_OPS_PER_CTU = 24020326        # see DESIGN.md 3.4 (sum over the 593 partitions)
CPUME_BIN = os.path.join(ROOT, "oracle", "_ref", "TAppEncoder_cpume")
CPUME_CFG = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_lowdelay_P_main.cfg")
CPUME_CFG_RA = os.path.join(ROOT, "oracle", "_ref", "cfg", "encoder_randomaccess_main.cfg")


def write_clip(path, frames):
    with open(path, "wb") as fh:
        for y in frames:
            h, w = y.shape
            fh.write(np.ascontiguousarray(y).tobytes())
            fh.write(np.full((h // 2) * (w // 2) * 2, 128, np.uint8).tobytes())


def reference_cpu_me(R, frames, procs, fast_search=0, cfg=None):
    """The reference's OWN CPU integer ME (--OpenCL=0 --FastSearch=0: TEncSearch::xPatternSearch + TComRdCost::xGetSAD*,
    TEncSearch.cpp:3774-3791,3835-3897) timed inside the reference encoder built from source with the counters of
    BASELINE.md section 3 (oracle/patch_cpume.py): `procs` independent single-threaded encoder processes (HM has no threads)
    encode the same 2-frame clip `frames` (I + P, one reference picture).  Returns block-SAD evaluations/s summed over
    processes, ME seconds (max over processes), DistFunc calls, wall seconds."""
    import re
    import tempfile
    H, W = frames[0].shape
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        write_clip(yuv, frames)
        cmds = [[CPUME_BIN, "-c", cfg or CPUME_CFG, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(len(frames)), "-q", "32",
                 "-b", os.path.join(d, "o%d.hevc" % i), "-o", "", "--OpenCL=0", "--FastSearch=%d" % fast_search, "--SearchRange=%d" % R] for i in range(procs)]
        t0 = time.perf_counter()
        ps = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for c in cmds]
        outs = [p.communicate()[0] for p in ps]
        wall = time.perf_counter() - t0
    secs, calls = [], []
    for o in outs:
        m = re.search(r"HMME_CPUME me_seconds=([0-9.]+) dist_calls=(\d+)", o)
        if not m:
            raise RuntimeError("reference encoder did not report:\n" + o[-1500:])
        secs.append(float(m.group(1)))
        calls.append(int(m.group(2)))
    return sum(calls) / max(secs), max(secs), sum(calls), wall


def reference_cpu_frac():
    """The reference's own xPatternSearchFracDIF (TEncSearch.cpp:4294-4331), timed inside the instrumented reference encoder
    (oracle/patch_cpume.py) on one core: default fast integer search (so the run is short), 416x240, I + P."""
    import re
    import tempfile
    W, H = 416, 240
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        write_clip(yuv, luma_frames(W, H, 2))
        r = subprocess.run([CPUME_BIN, "-c", CPUME_CFG, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "2", "-q", "32",
                            "-b", os.path.join(d, "o.hevc"), "-o", "", "--OpenCL=0", "--FastSearch=1", "--SearchRange=64"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    m = re.search(r"frac_seconds=([0-9.]+) frac_calls=(\d+) frac_pixels=(\d+)", r.stdout)
    if not m:
        raise RuntimeError("reference encoder did not report the fractional refinement:\n" + r.stdout[-1500:])
    secs, calls, px = float(m.group(1)), int(m.group(2)), int(m.group(3))
    return {"pu_refinements_per_s": calls / secs, "pu_pixels_per_s": px / secs, "cores": 1, "kind": "reference",
            "sample": "%d xPatternSearchFracDIF calls (%d PU pixels) in %.3f s inside the reference encoder, 416x240 I+P, one core" % (calls, px, secs)}


def workload_crop(name, clip, nframes=2):
    """The top-left clip[0] x clip[1] crop of the workload's own synthetic frames (frame 0 = reference, frame 1 = current, ...)."""
    W, H, _ = WORKLOADS[name]
    cw, ch = min(clip[0], W // 64 * 64), min(clip[1], H // 64 * 64)
    f = luma_frames(W, H, nframes)
    return [np.ascontiguousarray(y[:ch, :cw]) for y in f], (cw, ch)


def cpu_baseline_entry(name, clip, all_cores=True):
    """cpu_baseline object: the reference's CPU ME on a fixed crop of the workload when oracle/_ref holds its build, else the oracle port."""
    W, H, R = WORKLOADS[name]
    threads = os.cpu_count() or 1
    if os.path.exists(CPUME_BIN) and os.path.exists(CPUME_CFG):
        ra = name.endswith("_ra")                                   # config[2]: the random-access configuration (two lists, bi-prediction refinement), 3 pictures
        frames, (cw, ch) = workload_crop(name, clip, 3 if ra else 2)
        procs = threads if all_cores else 1
        v, me_s, calls, wall = reference_cpu_me(R, frames, procs, cfg=CPUME_CFG_RA if ra else None)
        return {"value": v, "unit": "block-SAD evaluations/s", "cores": procs, "kind": "reference",
                "sample": "reference encoder built from source (oracle/_ref/TAppEncoder_cpume), %s, --OpenCL=0 --FastSearch=0 --SearchRange=%d, top-left %dx%d crop "
                          "(%d CTUs) of the workload's frames (%s), %d independent single-threaded processes: %d DistFunc calls in %.1f s of "
                          "xPatternSearch (max over processes), %.1f s wall"
                          % ("encoder_randomaccess_main.cfg" if ra else "encoder_lowdelay_P_main.cfg", R, cw, ch, (cw // 64) * (ch // 64),
                             "3 pictures, two lists + bi-prediction refinement" if ra else "I+P, 1 ref", procs, calls, me_s, wall)}, wall
    v, dt, n = cpu_oracle_throughput(W, H, R, max(80, R + 16), max(8 * threads, 64), threads)
    return {"value": v, "unit": "block-SAD evaluations/s", "cores": threads, "kind": "port",
            "sample": "%d CTU jobs of the frame, %.1f s wall on %d threads (oracle/hmme_oracle.c; oracle/_ref absent)" % (n, dt, threads)}, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of integer ME on the host cores.  Every step times the SAME fixed
    sample (REF_ARM_CLIP crop of the workload's frames), whatever --steps is."""
    if rank != 0:
        return
    W, H, R, margin = workload_geometry(args.workload)
    # Each repeat is ~10 s of encoder processes on every host core.  The sample stays fixed; what shrinks when --steps is large is the
    # number of repeats actually timed: one warm-up repeat, then timed repeats until --steps of them are done or REF_ARM_BUDGET_S of wall
    # clock is used (at least two), so that the whole run ends within a few minutes whatever the driver passes.
    vals, entry = [], None
    t_begin = time.time()
    warm = min(args.warmup, 1)
    for s in range(warm + args.steps):
        entry, wall = cpu_baseline_entry(args.workload, REF_ARM_CLIP)
        if s >= warm:
            vals.append((entry["value"], wall))
        if len(vals) >= 2 and time.time() - t_begin + wall > REF_ARM_BUDGET_S:
            break
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals])) * 1e3
    entry["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": "me_block_sad_evaluations_per_s", "value": value, "unit": "block-SAD evaluations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "s16", "data": "synthetic",
        "frames_per_s_equivalent": value / (NPARTS * (2 * R + 1) ** 2 * (W // 64) * (H // 64)),
        "config": shared_config(args.workload),
        "repeats_timed": len(vals), "repeats_warmup": warm,
        "repeats_note": "every repeat times the same fixed sample; repeats beyond a %d s wall-clock budget are not run" % REF_ARM_BUDGET_S,
        "cpu_baseline": entry,
        "e2e": {"value": value, "unit": "block-SAD evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 200 (b200 arm), 3 (reference arm)")
    ap.add_argument("--warmup", type=int, default=None, help="default: 3 (b200 arm), 1 (reference arm)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p64", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the bit-exact comparison of every rank's band with the CPU oracle after the timed regions")
    ap.add_argument("--no-extras", action="store_true", help="skip the fractional-refinement / distortion / per-CTU / random-access legs (profiling runs)")
    ap.add_argument("--ref-dist", default="band_halo", choices=["band_halo", "broadcast"],
                    help="how the reference picture reaches the GPUs in the reported e2e leg (the other one is measured as well at N > 1)")
    ap.add_argument("--slots", type=int, default=0, choices=[0, 2, 3], help="frames in flight in the pipelined e2e leg (0 = hmme_group_pipeline_depth's advice)")
    ap.add_argument("--virtual-world", type=int, default=0, help="experiments: on ONE GPU, run only the band rank 0 would get in a world of this size")
    args = ap.parse_args()
    if args.impl == "reference":          # each step is seconds of single-threaded CPU encoders: keep the default run short
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
    else:
        args.steps = 200 if args.steps is None else args.steps
        args.warmup = max(3 if args.warmup is None else args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from _pkg import hm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        box = [hm.Group.unique_id() if rank == 0 else None]          # the library's own NCCL communicator: id made by rank 0, carried by the launcher's group
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    W, H, R, margin = workload_geometry(args.workload)
    ncx, ncy = W // 64, H // 64
    vworld = args.virtual_world or world
    all_jobs = frame_jobs(W, H, R)
    lib = hm.HmmeLib.get()
    if args.virtual_world:                                            # one GPU plays rank 0 of a larger world: only that band exists
        f0, n0 = lib.band_split(len(all_jobs), vworld, 0)
        all_jobs = np.ascontiguousarray(all_jobs[f0:f0 + n0])
    total_jobs = len(all_jobs)
    cands_per_job = (2 * R + 1) ** 2

    grp = hm.Group(device=local_rank, rank=rank, world=world, unique_id=uid, max_search_range=max(R, 64))
    grp.set_lambda_q16(LAMBDA_Q16)
    first, njobs = grp.band(total_jobs)
    jobs = np.ascontiguousarray(all_jobs[first:first + njobs])
    mes = [grp.context(0, s) for s in range(2)]                       # the per-GPU contexts behind two of the group's frame slots (resident legs)
    me = mes[0]
    exts = [torch.cuda.ExternalStream(m.stream_ptr, device=dev) for m in mes]

    # pinned host frames: HM's sample type (Pel = int16) and the same content as uint8; synthetic
    f = luma_frames(W, H, 2)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    n_cur, n_ref = pin(pad_plane(f[1], margin, margin)), pin(pad_plane(f[0], margin, margin))
    n_cur8, n_ref8 = pin(n_cur.astype(np.uint8)), pin(n_ref.astype(np.uint8))
    org = (margin, margin)
    NSLOTS = args.slots or grp.pipeline_depth(total_jobs, R)           # frames in flight in the pipelined e2e leg: the library's advice (2 or 3)
    outs = [[pin(np.zeros((total_jobs, NPARTS), t)) for t in (np.int32, np.int32, np.uint32, np.uint32)] for _ in range(NSLOTS)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    # ------------------------------------------------------------------ value: inputs resident in HBM
    # NSETS copies of the (current, reference) plane pair at distinct addresses, together larger than twice the 126 MB L2, cycled
    # through step by step ("inputs larger than L2"; the kernel is compute bound, but the rule is kept).
    plane_bytes = (H + 2 * margin) * ((W + 2 * margin + 15) // 16 * 16)
    nsets = max(2, -(-2 * 126 * (1 << 20) // (2 * plane_bytes)))
    sets = []
    for k in range(nsets):
        pc, pr = me.alloc_plane(1, W, H, margin, margin), me.alloc_plane(1, W, H, margin, margin)
        me.upload(pc, n_cur8, asynchronous=True)
        me.upload(pr, n_ref8, asynchronous=True)
        sets.append((pc, pr))
    me.sync()
    barrier()
    peak = me.measure_int_alu_peak()

    # (a) dominant-kernel duration and single-stream step time: one context, CUDA events per step on its stream
    for s in range(args.warmup):
        if njobs:
            me.search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
    barrier()
    n_single = min(args.steps, 40)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_single)]
    kernel_ms = []
    for s in range(n_single):
        ev[s][0].record(exts[0])
        if njobs:
            me.search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
        ev[s][1].record(exts[0])
        if njobs:
            kernel_ms.append(me.last_kernel_ms())                  # CUDA events around the dominant kernel, same stream
    barrier()
    single_ms = sum(a.elapsed_time(b) for a, b in ev) / n_single

    # (b) the reported value: EXACTLY K steps, frames alternating over two slots' contexts (two streams), nothing but the
    # library's kernels in the timed region; consecutive frames overlap at their wave tails.  CUDA events around the whole region.
    for s in range(max(args.warmup, 4)):
        if njobs:
            mes[s & 1].search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = grp.kernel_launches
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), [torch.cuda.Event(enable_timing=True) for _ in mes]
    barrier()
    wall0 = time.perf_counter()
    ev_start.record(exts[0])
    for s in range(args.steps):
        if njobs:
            mes[s & 1].search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
    for x_, e_ in zip(exts, ev_end):
        e_.record(x_)
    barrier()
    wall1 = time.perf_counter()
    launches = grp.kernel_launches - launches0
    total_ms, single_ms, kern_ms = allmax([max(ev_start.elapsed_time(e_) for e_ in ev_end), single_ms, float(np.mean(kernel_ms)) if kernel_ms else 0.0])
    clocks = sampler.summary() if sampler else None
    resident = None
    if njobs and not args.no_verify:
        resident = me.fetch_results(njobs)                         # the last resident search of context 0: verified below with the e2e results

    # ------------------------------------------------------------------ e2e: host planes through the group call (C++ does the band split,
    # the rectangle uploads / NCCL broadcast, the search and the result copies; Python issues one call per frame)
    # (a) serial: every step waits for its own results before the next upload starts (a low-delay encoder's dependency);
    # (b) pipelined (the reported e2e): frames cycle through the group's three slots, so the copies of one frame overlap the kernels of
    #     the others -- every step still uploads both planes' rectangles and downloads its four result arrays.
    def e2e_leg(ref_dist, cur_h, ref_h, steps):
        grp.configure(W, H, margin, margin, grp.BROADCAST if ref_dist == "broadcast" else grp.BAND_HALO)
        calls = [grp.bind_frame(s, cur_h, org, ref_h, org, all_jobs, R, outs[s]) for s in range(NSLOTS)]
        syncs = [grp.bind_sync(s) for s in range(NSLOTS)]
        for s in range(2 * NSLOTS):
            calls[s % NSLOTS]()
            syncs[s % NSLOTS]()
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            calls[0]()
            syncs[0]()
        barrier()
        serial = (time.perf_counter() - t0) * 1e3 / steps
        t0 = time.perf_counter()
        for s in range(steps):
            syncs[s % NSLOTS]()                                     # this slot's previous frame (NSLOTS steps ago) is complete
            calls[s % NSLOTS]()
        for sy in syncs:
            sy()
        barrier()
        piped = (time.perf_counter() - t0) * 1e3 / steps
        piped, serial = allmax([piped, serial])
        return piped, serial

    e2e = {}
    legs = [(args.ref_dist, "s16")]
    if not args.no_extras:
        legs.append((args.ref_dist, "u8"))
        if world > 1:
            legs.append(("broadcast" if args.ref_dist == "band_halo" else "band_halo", "s16"))
    for ref_dist, ty in legs[::-1]:                                 # the reported leg last: its results stay in `outs` for the verification
        e2e[(ref_dist, ty)] = e2e_leg(ref_dist, n_cur if ty == "s16" else n_cur8, n_ref if ty == "s16" else n_ref8, args.steps)
    e2e_ms, serial_ms = e2e[legs[0]]
    if njobs:
        cr, rr = lib.band_extent(jobs, R)
    else:
        cr = rr = (0, 0, 0, 0)
    rect_px = lambda r: (r[2] - r[0]) * (r[3] - r[1])              # noqa: E731
    ref_px = ((W + 2 * margin) * (H + 2 * margin) if rank == 0 else 0) if args.ref_dist == "broadcast" and world > 1 else rect_px(rr)
    h2d, d2h = allsum([2 * (ref_px + rect_px(cr)) + jobs.nbytes, 4 * njobs * NPARTS * 4])

    # ------------------------------------------------------------------ verification: EVERY CTU of this rank's band, bit for bit, against the CPU oracle
    verified = None
    if not args.no_verify:
        from oracle.pyoracle import Oracle                          # the checker, after the timed regions; never the thing measured
        mism, checked = 0, 0
        if njobs:
            want = Oracle().search_frame(n_cur, org, n_ref, org, jobs, R, LAMBDA_Q16, nthreads=max(1, (os.cpu_count() or 1) // max(1, world)))
            for got in [[o[first:first + njobs] for o in outs[s]] for s in range(NSLOTS)] + [resident]:
                mism += int(sum(int((np.asarray(g) != np.asarray(w_)).any(axis=1).sum()) for g, w_ in zip(got, want)))
                checked += njobs
        mism, checked, band = allsum([mism, checked, njobs])
        verified = {"ctus": int(band), "ctu_result_sets_compared": int(checked), "mismatches": int(mism),
                    "what": "X, Y, sad and cost of all 593 partitions of every CTU job of every rank's band (results of every e2e frame slot and of the last "
                            "resident search) against oracle.search_frame on the same inputs"}

    # ------------------------------------------------------------------ next row: fractional-pel refinement
    frac = None
    if not args.virtual_world and not args.no_extras:
        nfr = min(args.steps, 60)
        fk, fk_sad = [], []
        for s in range(nfr + 2 if njobs else 0):           # kernel time: CUDA events inside the library, resident inputs
            pc_, pr_ = sets[s % nsets]
            me.search_frame_async(pc_, pr_, jobs, R)
            me.refine_frame(pc_, pr_, njobs, None, True, asynchronous=True)
            me.sync()
            if s >= 2:
                fk.append(me.last_frac_ms())
        for s in range(6 if njobs else 0):
            pc_, pr_ = sets[s % nsets]
            me.search_frame_async(pc_, pr_, jobs, R)
            me.refine_frame(pc_, pr_, njobs, None, False, asynchronous=True)
            me.sync()
            fk_sad.append(me.last_frac_ms())
        if not njobs:
            fk, fk_sad = [0.0], [0.0]
        barrier()
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record(exts[0])
        for s in range(nfr):                                # search + refinement per frame, one context: the two kernels of a frame
            pc_, pr_ = sets[s % nsets]                      # depend on each other and both fill the GPU, so there is nothing to overlap
            if njobs:
                me.search_frame_async(pc_, pr_, jobs, R)
                me.refine_frame(pc_, pr_, njobs, None, True, asynchronous=True)
        ev_b.record(exts[0])
        barrier()
        both_ms = ev_a.elapsed_time(ev_b) / nfr
        fkm, fksm, both_ms = allmax([float(np.mean(fk)), float(np.mean(fk_sad)), both_ms])
        pu_px = total_jobs * 24 * 4096                      # sum of the 593 partition areas = 24 CTU areas
        # algorithmic operations of the reference's own scheme per CTU (all 593 partitions, half-pel winner at the centre): filter MACs over
        # the plane sizes of xExtDIFUpSamplingH/Q + 8 (8x8 Hadamard) or 6 (4x4) operations per pixel and candidate; formula in DESIGN.md 3.4
        frac_ops = FRAC_OPS_PER_CTU * njobs
        ach = frac_ops / (fkm * 1e-3) if fkm > 0 else 0.0
        frac = {"scope": "fractional-pel refinement (xPatternSearchFracDIF: 9 half-pel + 9 quarter-pel candidates, 8-tap interpolation, Hadamard cost) "
                         "of all 593 partitions of every CTU, from the integer winners left on the device",
                "kernel": "me_frac_group_kernel", "pus_per_frame": total_jobs * NPARTS, "kernel_ms": fkm, "kernel_ms_sad": fksm,
                "pu_refinements_per_s": total_jobs * NPARTS / (fkm * 1e-3) if fkm else None, "pu_pixels_per_s": pu_px / (fkm * 1e-3) if fkm else None,
                "search_plus_refine_ms_per_frame": both_ms, "search_plus_refine_frames_per_s": 1e3 / both_ms if both_ms else None, "steps": nfr,
                "roofline": {"bound": "int_issue", "achieved": ach / 1e12, "peak": 2.0 * peak["lane_ops_per_s"] / 1e12, "unit": "T int-op/s",
                             "frac": ach / (2.0 * peak["lane_ops_per_s"]) if peak["lane_ops_per_s"] else None,
                             "frac_issue": ach / (2.0 * peak["lane_ops_per_s"]) if peak["lane_ops_per_s"] else None,
                             "frac_one_pipe": ach / peak["lane_ops_per_s"] if peak["lane_ops_per_s"] else None,
                             "ops_per_ctu": FRAC_OPS_PER_CTU,
                             "peak_source": "peak = both integer pipes (ALU + FMA-heavy/IMAD, 2 x the live-measured 64 lanes/clk/SM = every issue slot), the same ceiling as the "
                                            "search kernel's; frac = frac_issue = achieved / peak; frac_one_pipe = against one pipe (round 1's denominator)"},
                "timer": "kernel_ms: CUDA events around me_frac_group_kernel on its stream; search_plus_refine: CUDA events around K frames on one context/stream, "
                         "resident inputs"}

    # ------------------------------------------------------------------ row f3: motion-compensated distortion at quarter-pel MVs
    mc = None
    if world == 1 and njobs and not args.virtual_world and not args.no_extras:
        rng = np.random.default_rng(7)
        rects = lib.partition_table()
        mpus = np.zeros((njobs, NPARTS, 6), np.int32)
        mpus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
        mpus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
        mpus[:, :, 2], mpus[:, :, 3] = rects[None, :, 2], rects[None, :, 3]
        mpus[:, :, 4:6] = rng.integers(-4 * (R - 8), 4 * (R - 8), size=(njobs, NPARTS, 2))     # any quarter-pel MV inside the search range
        km = {}
        for name, had in (("sad", False), ("hadamard", True)):
            t = []
            for s in range(4):
                me.mc_cost(sets[s % nsets][0], sets[s % nsets][1], mpus.reshape(-1, 6), had)
                t.append(me.last_frac_ms())
            km[name] = float(np.mean(t[1:]))
        bpus = np.concatenate([mpus, rng.integers(-4 * (R - 8), 4 * (R - 8), size=(njobs, NPARTS, 2)).astype(np.int32)], axis=2).reshape(-1, 8)
        tb = []
        for s in range(3):                                   # bi-directional: the reference plane and the current plane stand in for the two lists
            me.mc_cost_bi(sets[s % nsets][0], sets[s % nsets][1], sets[s % nsets][0], bpus, True)
            tb.append(me.last_frac_ms())
        km["bi_hadamard"] = float(np.mean(tb[1:]))
        mc = {"scope": "distortion of the motion-compensated uni-prediction (8-tap interpolation at a quarter-pel MV) of all 593 partitions of every CTU: "
                       "the arithmetic of xGetTemplateCost (SAD) / uni-directional merge candidates (Hadamard)",
              "kernel": "me_mc_group_kernel", "pus": njobs * NPARTS, "kernel_ms_sad": km["sad"], "kernel_ms_hadamard": km["hadamard"], "kernel_ms_bi_hadamard": km["bi_hadamard"],
              "pu_pixels_per_s_sad": njobs * 24 * 4096 / (km["sad"] * 1e-3), "timer": "CUDA events around the kernel on its stream, resident planes"}

    # ------------------------------------------------------------------ the call the encoder makes: synchronous per-CTU search with HOST pointers
    per_ctu = None
    if world == 1 and njobs and not args.virtual_world and not args.no_extras:
        per_ctu = per_ctu_leg(hm, me, n_cur, n_ref, margin, all_jobs, R, W, H)

    # ------------------------------------------------------------------ BASELINE config[2]: random access, two lists + bi-prediction refinement
    ra = None
    if args.workload == "1080p64_ra" and world == 1 and not args.no_extras:
        ra = random_access_leg(hm, me, exts[0], torch, sets, nsets, n_cur, n_ref, f, margin, all_jobs, R, W, H, peak, args)

    rc = 0
    if rank == 0:
        total_cands = total_jobs * cands_per_job
        ms_per_step = total_ms / args.steps
        value = total_cands * NPARTS / (ms_per_step * 1e-3)
        e2e_value = total_cands * NPARTS / (e2e_ms * 1e-3)
        # dominant kernel roofline: algorithmic integer lane-ops of THIS rank's launch / its CUDA-event duration
        cands_rank = njobs * cands_per_job
        achieved = cands_rank * INT_OPS_PER_CAND / (kern_ms * 1e-3) if kern_ms > 0 else 0.0
        alg_bytes = (W + 2 * margin) * (H + 2 * margin) + W * H + total_jobs * NPARTS * 16
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        cfg = shared_config(args.workload)
        out = {
            "metric": "me_block_sad_evaluations_per_s", "value": value, "unit": "block-SAD evaluations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "frames_per_s": 1e3 / ms_per_step,
            "ctu_candidates_per_s": total_cands / (ms_per_step * 1e-3),
            "config": cfg,
            "detail": {"sharding": "CTU-row bands cut at CTU granularity by the library (hmme_band_split) over %d GPU(s), one process per GPU; reference picture: %s"
                                   % (world, "single GPU" if world == 1 else ("band + halo rectangle per GPU over its own PCIe link, no collective" if args.ref_dist == "band_halo"
                                                                              else "rank 0 uploads, ncclBroadcast over NVLink inside the library")),
                       "l2": "inputs larger than L2: %d resident (current, reference) plane pairs at distinct addresses (%.0f MiB), cycled step by step" % (nsets, 2 * nsets * plane_bytes / 2**20),
                       "timer": "CUDA events around the whole K-step region, frames alternating over two of the group's frame slots (two contexts/streams), max over ranks; "
                                "single_stream_ms_per_step = one context, events per step"},
            "single_stream_ms_per_step": single_ms,
            "clocks": clocks,
            "gpu_launches": int(launches),
            "wall_ms_timed_region": (wall1 - wall0) * 1e3,
            "e2e": {"value": e2e_value, "unit": "block-SAD evaluations/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "frames_per_s": 1e3 / e2e_ms, "ms_per_step": e2e_ms,
                    "serial_ms_per_step": serial_ms, "serial_frames_per_s": 1e3 / serial_ms, "slots": NSLOTS,
                    "host_samples": "int16 (HM's Pel), page-locked; narrowed to 8 bit on the device", "ref_dist": args.ref_dist,
                    "timer": "host wall clock around K x hmme_group_search_frame_async (+ hmme_group_sync of the slot `slots` steps back): per rank, rectangle uploads of the band's rows of "
                             "the current frame and of band + halo of the reference picture from pinned memory, jobs, search, four result arrays back; frames cycle through 2 or 3 of the group's slots (hmme_group_pipeline_depth; `slots` below) "
                             "so copies overlap kernels; serial_* = one slot, each step waits for its results; max over ranks"},
            "roofline": {"bound": "int_issue", "kernel": "me_u8_tile_kernel", "achieved": achieved / 1e12, "peak": 2.0 * peak["lane_ops_per_s"] / 1e12,
                         "unit": "T int-lane-op/s", "frac": achieved / (2.0 * peak["lane_ops_per_s"]) if peak["lane_ops_per_s"] else None,
                         "frac_issue": achieved / (2.0 * peak["lane_ops_per_s"]) if peak["lane_ops_per_s"] else None,
                         "frac_one_pipe": achieved / peak["lane_ops_per_s"] if peak["lane_ops_per_s"] else None,
                         "traffic": _ncu_traffic(args.workload) if world == 1 else None,
                         "ops_per_ctu_candidate": INT_OPS_PER_CAND, "kernel_ms": kern_ms,
                         "frac_at_step_rate": (cands_rank * INT_OPS_PER_CAND / (ms_per_step * 1e-3)) / (2.0 * peak["lane_ops_per_s"]) if peak["lane_ops_per_s"] else None,
                         "pixel_abs_diffs_per_s": cands_rank * PX_PER_CAND / (kern_ms * 1e-3) if kern_ms > 0 else 0.0,
                         "peak_source": "measured live: VABSDIFF4.U8.ACC issue rate, %.1f lanes/clk/SM x %d SMs at %.0f MHz = ONE integer pipe; `peak` = both integer pipes "
                                        "(ALU + FMA-heavy/IMAD = 2 x that = 128 lanes/clk/SM, i.e. every issue slot), the same ceiling for every kernel of this line, and `frac` = "
                                        "`frac_issue` = achieved / peak.  `frac_one_pipe` keeps round 1's denominator (the survey's: one pipe); it exceeds 1 since the kernel "
                                        "needs fewer instructions than the survey's 2803 operations per CTU-candidate (key algebra) and runs its additions on the second pipe, so "
                                        "one pipe is not a ceiling for that count"
                                        % (peak["lanes_per_clk_sm"], sms, peak["sm_mhz"]),
                         "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0,
                                 "peak_gbs": _measured_hbm()}},
        }
        for (rd, ty), (p_ms, s_ms) in e2e.items():
            if (rd, ty) != legs[0]:
                out["e2e"]["%s_%s" % (rd, ty)] = {"ms_per_step": p_ms, "serial_ms_per_step": s_ms, "value": total_cands * NPARTS / (p_ms * 1e-3)}
        if verified is not None:
            out["verified"] = verified
            if verified["mismatches"] or verified["ctus"] != total_jobs:
                rc = 1
        if frac:
            out["frac_refine"] = frac
        if mc:
            out["mc_cost"] = mc
        if per_ctu:
            out["per_ctu"] = per_ctu
        if ra:
            out["random_access"] = ra
            if ra.get("verified", {}).get("mismatches"):
                rc = 1
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], _ = cpu_baseline_entry(args.workload, (416, 240))
            if frac and os.path.exists(CPUME_BIN):
                fr = reference_cpu_frac()
                frac["cpu_reference"] = fr
                frac["gpu_over_one_core"] = frac["pu_pixels_per_s"] / fr["pu_pixels_per_s"]
            if os.path.exists(CPUME_BIN):       # the reference's default (fast) integer search, for context: TZ evaluates ~600x fewer candidates
                v, me_s, calls, wall = reference_cpu_me(R, luma_frames(416, 240, 2), 1, fast_search=1)
                out["cpu_tz"] = {"value": v, "unit": "block-SAD evaluations/s", "cores": 1, "kind": "reference",
                                 "sample": "same encoder, --FastSearch=1 (xTZSearch): %d DistFunc calls in %.3f s of integer ME for the P frame of a "
                                           "416x240 clip (18 full CTUs) on one core" % (calls, me_s)}
            threads = os.cpu_count() or 1
            v, dt, n = cpu_oracle_throughput(W, H, R, margin, max(8 * threads, 64), threads)
            out["cpu_port"] = {"value": v, "unit": "block-SAD evaluations/s", "cores": threads, "kind": "port",
                               "sample": "GPU-ME semantics on the CPU (oracle/hmme_oracle.c): %d of %d CTU jobs of the same frame, %.1f s wall on %d threads" % (n, total_jobs, dt, threads)}
        print(json.dumps(out))
    for pc, pr in sets:
        pc.free()
        pr.free()
    grp.close()
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit("bench.py: verification against the oracle FAILED (see the `verified` object of the JSON line)")


def per_ctu_leg(hm, me, n_cur, n_ref, margin, all_jobs, R, W, H):
    """hmme_search_ctu as TEncSearch::xMotionEstimation calls it (TEncSearch.cpp:3749): one CTU, host pointers into the int16 planes,
    synchronous.  Latency per call over every CTU of the frame, and the frame rate 480 sequential calls give."""
    import ctypes as C
    L = me.lib.L
    stride = n_ref.shape[1]
    X, Y = np.zeros(NPARTS, np.int32), np.zeros(NPARTS, np.int32)
    S, Cs = np.zeros(NPARTS, np.uint32), np.zeros(NPARTS, np.uint32)
    blocks = [np.ascontiguousarray(n_cur[margin + j[1]:margin + j[1] + 64, margin + j[0]:margin + j[0] + 64]) for j in all_jobs]

    def call(k):
        j = all_jobs[k]
        off = int(((margin + j[1]) * stride + margin + j[0]) * 2)
        rc = L.hmme_search_ctu(me.h, blocks[k].ctypes.data, 64, n_ref.ctypes.data + off, stride, int(R), int(j[2]), int(j[3]),
                               X.ctypes.data, Y.ctypes.data, S.ctypes.data, Cs.ctypes.data)
        if rc:
            raise hm.HmmeError(rc, L.hmme_last_error(me.h).decode())

    for k in range(min(8, len(all_jobs))):
        call(k)
    t0 = time.perf_counter()
    for k in range(len(all_jobs)):
        call(k)
    dt = time.perf_counter() - t0
    kms = me.last_kernel_ms()
    spec = None
    probe = os.path.join(ROOT, "tools", "spec_probe")
    if os.path.exists(probe):                                       # the C++ class over a whole picture: synchronous calls vs the speculative whole-frame search
        r = subprocess.run([probe, str(W), str(H), str(R)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and line:
            p = json.loads(line[-1])
            spec = {"what": "TEncOpenCL (C++ drop-in class) over one picture, one calcMotionVectors call per CTU in coding order: `sync` = every call searches synchronously; "
                            "`spec_hit` = beginPicture/addReferencePicture/speculate first, zero predictors, every call answered from the device-resident tables; "
                            "`spec_drift` = the predictor changes every 40 CTUs (miss -> synchronous search + re-speculation of the CTUs still to come, then hits)",
                    "sync_frame_ms": p["sync_frame_ms"], "sync_frames_per_s": 1e3 / p["sync_frame_ms"],
                    "spec_hit_frame_ms": p["spec_hit_frame_ms"], "spec_hit_frames_per_s": 1e3 / p["spec_hit_frame_ms"],
                    "spec_drift_frame_ms": p["spec_drift_frame_ms"], "spec_drift_frames_per_s": 1e3 / p["spec_drift_frame_ms"],
                    "spec_hit_rate": p["spec_hit_hits"] / max(1, p["spec_hit_calls"]), "spec_drift_hit_rate": p["spec_drift_hits"] / max(1, p["spec_drift_calls"]),
                    "tables_equal_sync": p["tables_equal_sync"], "slowest_of_4_ms": p.get("slowest_ms"),
                    "timer": "host wall clock inside tools/spec_probe (includes the picture uploads from pageable memory); fastest of 4 timed pictures per mode"}
        else:
            spec = {"error": (r.stderr or r.stdout)[-300:]}
    return {"speculative": spec, "call": "hmme_search_ctu (TEncOpenCL::calcMotionVectors): one CTU, +-%d, host pointers into int16 planes, synchronous" % R,
            "calls": len(all_jobs), "latency_ms": dt * 1e3 / len(all_jobs), "frame_ms": dt * 1e3, "frames_per_s": 1.0 / dt,
            "search_kernel_ms_last_call": kms,
            "timer": "host wall clock around %d sequential calls (every CTU of the frame), after 8 warm-up calls" % len(all_jobs)}


def random_access_leg(hm, me, ext, torch, sets, nsets, n_cur, n_ref, frames, margin, all_jobs, R, W, H, peak, args):
    """BASELINE config[2]: one B frame = two reference lists (one +-R search each, 8-bit) + the bi-prediction refinement
    (TEncSearch.cpp:3168-3221: the block becomes 2*org - pred of the other list, 16 bit, searched +-4 around the list's winner,
    :3702-3712), all three result sets kept in a device-resident table (slots 0/1/2)."""
    from oracle.pyoracle import Oracle
    njobs = len(all_jobs)
    # list 1 reference: the frame after the current one (the synthetic pan continues); bi-prediction block: 2*org - list-0 reference
    f3 = luma_frames(W, H, 3)
    n_ref1 = pad_plane(f3[2], margin, margin)
    n_bi = (2 * n_cur.astype(np.int32) - n_ref.astype(np.int32)).astype(np.int16)
    p_ref1, p_bi = me.alloc_plane(1, W, H, margin, margin), me.alloc_plane(2, W, H, margin, margin)
    me.upload(p_ref1, n_ref1)
    me.upload(p_bi, n_bi)
    table = me.create_table(3, njobs)
    pc, pr = sets[0]
    me.search_frame_table(pc, p_ref1, all_jobs, R, table, 1)
    x1, y1, _, _ = me.table_fetch(table, 1, 0, njobs)
    bi_jobs = all_jobs.copy()
    bi_jobs[:, 2] = x1[:, 592] - BI_RANGE                            # window centred on the 64x64 winner of list 1
    bi_jobs[:, 3] = y1[:, 592] - BI_RANGE

    def frame():
        me.search_frame_table(pc, pr, all_jobs, R, table, 0)
        me.search_frame_table(pc, p_ref1, all_jobs, R, table, 1)
        me.search_frame_table(p_bi, p_ref1, bi_jobs, BI_RANGE, table, 2)

    for _ in range(3):
        frame()
    me.sync()
    bi_ms = []
    for _ in range(5):
        me.search_frame_table(p_bi, p_ref1, bi_jobs, BI_RANGE, table, 2)
        me.sync()
        bi_ms.append(me.last_kernel_ms())
    n = min(args.steps, 50)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(ext)
    for _ in range(n):
        frame()
    b.record(ext)
    me.sync()
    ms = a.elapsed_time(b) / n
    res = [me.table_fetch(table, s, 0, njobs) for s in range(3)]
    ver = None
    if not args.no_verify:
        o = Oracle()
        th = os.cpu_count() or 1
        want = [o.search_frame(n_cur, (margin, margin), n_ref, (margin, margin), all_jobs, R, LAMBDA_Q16, nthreads=th),
                o.search_frame(n_cur, (margin, margin), n_ref1, (margin, margin), all_jobs, R, LAMBDA_Q16, nthreads=th),
                o.search_frame(n_bi, (margin, margin), n_ref1, (margin, margin), bi_jobs, BI_RANGE, LAMBDA_Q16, nthreads=th)]
        mism = sum(int((np.asarray(g) != np.asarray(w_)).any(axis=1).sum()) for got, wnt in zip(res, want) for g, w_ in zip(got, wnt))
        ver = {"ctus": 3 * njobs, "mismatches": int(mism), "what": "list 0, list 1 and bi-prediction result tables of every CTU against the oracle"}
    bi = float(np.mean(bi_ms[1:]))
    cands = njobs * (2 * BI_RANGE + 1) ** 2
    ach = cands * INT_OPS_PER_CAND_16 / (bi * 1e-3)
    me.destroy_table(table)
    p_ref1.free()
    p_bi.free()
    return {"scope": "one B frame of 1080p random access: list 0 and list 1 (480 jobs x 16641 candidates each, 8 bit) + bi-prediction refinement "
                     "(480 jobs x 81 candidates, int16 block 2*org-pred against the 8-bit list-1 picture), results in a device-resident [3][480][593] table",
            "ms_per_b_frame": ms, "b_frames_per_s": 1e3 / ms, "bipred_kernel": "me_bipred_prep_kernel + me_u8_tile_kernel (clamped block + per-partition constants, DESIGN.md 3.2)", "bipred_kernel_ms": bi,
            "bipred_roofline": {"bound": "int_issue", "achieved": ach / 1e12, "peak": 2 * peak["lane_ops_per_s"] / 1e12, "unit": "T int-lane-op/s",
                                "frac": ach / (2 * peak["lane_ops_per_s"]), "frac_issue": ach / (2 * peak["lane_ops_per_s"]), "frac_one_pipe": ach / peak["lane_ops_per_s"],
                                "ops_per_ctu_candidate": INT_OPS_PER_CAND_16},
            "block_sad_evaluations_per_s": (2 * njobs * (2 * R + 1) ** 2 + cands) * NPARTS / (ms * 1e-3),
            "timer": "CUDA events around %d B frames on one context/stream, resident planes" % n, "verified": ver}


def _ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture."""
    for name in ("r02d_traffic.json", "r02b_traffic.json", "r02_traffic.json", "r01_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            if t["workload"].startswith(workload.split("_")[0]):
                return t["traffic_bytes_per_launch"]
        except Exception:
            pass
    return None


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0   # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
