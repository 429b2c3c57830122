#!/usr/bin/env python
"""bench.py -- whole-CTU integer-pel motion estimation throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p64|4k128|...] [--impl reference]

A step = the integer ME of ONE frame against ONE reference picture: every full 64x64 CTU of the frame,
(2R+1)^2 candidates each, 593 partitions per candidate (config[1] of BASELINE.json at N=1: 1920x1080,
+-64 -> 480 jobs x 16641 candidates).  With N > 1 (torchrun, one rank per GPU) the frame is split into
contiguous CTU-row bands (strong scaling); the reference picture is uploaded by rank 0
and NCCL-broadcast, each rank uploads and searches only its band.

Output: ONE JSON line on rank 0 (contract in the task statement): `value` = block-SAD evaluations/s with
inputs resident in HBM (K frames alternating over two library contexts/streams, CUDA events around the whole region,
inputs cycled through more plane copies than fit in L2), `e2e` = the same
metric through the public C-ABI calls with pinned HOST buffers (H2D of both int16 planes + jobs, D2H of the
four result arrays inside the timed region), `roofline` = algorithmic integer lane-ops/s of the dominant
kernel against the ALU issue rate measured live, `cpu_baseline` = the reference's own CPU full-search ME
(oracle/_ref/TAppEncoder_cpume, one single-threaded process per core), `cpu_port` = the CPU oracle port.
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
    "4k128": (3840, 2160, 128),    # config[3]
    "1080p32": (1920, 1080, 32),
    "1080p16": (1920, 1080, 16),
    "416x240_64": (416, 240, 64),  # config[0] geometry
}
NPARTS = 593
INT_OPS_PER_CAND = 2803            # 1024 packed SADs + 593 adds + 593 cost adds + 593 min
PX_PER_CAND = 4096
LAMBDA_Q16 = 460000                # QP ~32 (SURVEY.md section 8d synthetic inputs)


def workload_geometry(name):
    W, H, R = WORKLOADS[name]
    margin = max(80, R + 16)       # HM pads by 80 (TComPicYuv.cpp:93-94); larger ranges need more for in-bounds windows
    return W, H, R, margin


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
        time.sleep(0.15)                       # let the first samples arrive before the region starts
        self.t0 = time.perf_counter()

    def summary(self):
        t1 = time.perf_counter()
        time.sleep(0.05)
        if self.proc:
            self.proc.terminate()
        self.thread.join(timeout=3)
        rows = [r for r, t in zip(self.rows, self.t) if self.t0 <= t <= t1 + 0.03] or self.rows[-3:]
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


def reference_cpu_me(R, clip, procs, fast_search=0):
    """The reference's OWN CPU integer ME (--OpenCL=0 --FastSearch=0: TEncSearch::xPatternSearch + TComRdCost::xGetSAD*,
    TEncSearch.cpp:3774-3791,3835-3897) timed inside the reference encoder built from source with the counters of
    BASELINE.md section 3 (oracle/patch_cpume.py): `procs` independent single-threaded encoder processes (HM has no threads)
    encode the same 2-frame synthetic clip (I + P, one reference picture).  Returns block-SAD evaluations/s summed over
    processes, ME seconds (max over processes), DistFunc calls."""
    import re
    import tempfile
    W, H = clip
    with tempfile.TemporaryDirectory() as d:
        yuv = os.path.join(d, "clip.yuv")
        with open(yuv, "wb") as fh:
            for y in luma_frames(W, H, 2):
                fh.write(y.tobytes())
                fh.write(np.full((H // 2) * (W // 2) * 2, 128, np.uint8).tobytes())
        cmds = [[CPUME_BIN, "-c", CPUME_CFG, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "2", "-q", "32",
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
        with open(yuv, "wb") as fh:
            for y in luma_frames(W, H, 2):
                fh.write(y.tobytes())
                fh.write(np.full((H // 2) * (W // 2) * 2, 128, np.uint8).tobytes())
        r = subprocess.run([CPUME_BIN, "-c", CPUME_CFG, "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", "2", "-q", "32",
                            "-b", os.path.join(d, "o.hevc"), "-o", "", "--OpenCL=0", "--FastSearch=1", "--SearchRange=64"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    m = re.search(r"frac_seconds=([0-9.]+) frac_calls=(\d+) frac_pixels=(\d+)", r.stdout)
    if not m:
        raise RuntimeError("reference encoder did not report the fractional refinement:\n" + r.stdout[-1500:])
    secs, calls, px = float(m.group(1)), int(m.group(2)), int(m.group(3))
    return {"pu_refinements_per_s": calls / secs, "pu_pixels_per_s": px / secs, "cores": 1, "kind": "reference",
            "sample": "%d xPatternSearchFracDIF calls (%d PU pixels) in %.3f s inside the reference encoder, 416x240 I+P, one core" % (calls, px, secs)}


def cpu_baseline_entry(R, budget_s, all_cores=True):
    """cpu_baseline object: the reference's CPU ME when oracle/_ref holds its build, else the oracle port."""
    threads = os.cpu_count() or 1
    if os.path.exists(CPUME_BIN) and os.path.exists(CPUME_CFG):
        clip = (416, 240) if budget_s >= 30 else (192, 128) if budget_s >= 10 else (128, 64) if budget_s >= 4 else (64, 64)
        procs = threads if all_cores else 1
        v, me_s, calls, wall = reference_cpu_me(R, clip, procs)
        return {"value": v, "unit": "block-SAD evaluations/s", "cores": procs, "kind": "reference",
                "sample": "reference encoder built from source (oracle/_ref/TAppEncoder_cpume), --OpenCL=0 --FastSearch=0 --SearchRange=%d, "
                          "%dx%d synthetic clip, 2 frames (I+P, 1 ref), %d independent single-threaded processes: %d DistFunc calls in %.1f s of "
                          "xPatternSearch (max over processes), %.1f s wall" % (R, clip[0], clip[1], procs, calls, me_s, wall)}, wall
    W, H, R_, margin = workload_geometry("1080p64")
    v, dt, n = cpu_oracle_throughput(W, H, R, max(80, R + 16), max(8 * threads, 64), threads)
    return {"value": v, "unit": "block-SAD evaluations/s", "cores": threads, "kind": "port",
            "sample": "%d CTU jobs of the 1080p frame, %.1f s wall on %d threads (oracle/hmme_oracle.c; oracle/_ref absent)" % (n, dt, threads)}, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of integer ME on the host cores (see reference_cpu_me)."""
    if rank != 0:
        return
    W, H, R, margin = workload_geometry(args.workload)
    budget = 240.0 / max(1, args.steps + args.warmup)
    vals, entry = [], None
    for s in range(args.warmup + args.steps):
        entry, wall = cpu_baseline_entry(R, budget)
        if s >= args.warmup:
            vals.append((entry["value"], wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals])) * 1e3
    entry["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": "me_block_sad_evaluations_per_s", "value": value, "unit": "block-SAD evaluations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "s16", "data": "synthetic",
        "frames_per_s_equivalent": value / (NPARTS * (2 * R + 1) ** 2 * (W // 64) * (H // 64)),
        "config": {"workload": "%s: %dx%d luma, 64x64 CTU, integer-pel full search +-%d, 1 reference picture" % (args.workload, W, H, R),
                   "sample": entry["sample"]},
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
    ap.add_argument("--ref-dist", default="broadcast", choices=["allgather", "broadcast"],
                    help="N > 1: how the reference plane reaches every GPU: rank 0 uploads all of it and NCCL broadcasts (default), or each "
                         "rank uploads 1/N of it and NCCL all-gathers (faster when steps are serial, no gain once frames are pipelined)")
    ap.add_argument("--graphs", action="store_true", help="e2e: replay CUDA graphs of the step (hmme_graph_*) instead of issuing it call by call; measured: "
                                                          "same at N = 1 (1.269 ms), 0.187 vs 0.191 ms for an 8-GPU-sized band on one GPU, slower at N = 2 "
                                                          "(0.74 vs 0.66 ms: the split around the NCCL broadcast loses overlap), hence off by default")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL kernels on a high-priority stream: the reference-plane broadcast of one context slips in between the search CTAs of the
        # other context instead of queueing behind them (HMME_NCCL_LOW_PRIORITY=1 restores the default for comparison)
        pg_opts = None
        if not os.environ.get("HMME_NCCL_LOW_PRIORITY"):
            try:
                pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:
                pg_opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)

    W, H, R, margin = workload_geometry(args.workload)
    ncx, ncy = W // 64, H // 64
    jobs, (r0, r1) = hm.band_jobs(W, H, R, args.virtual_world or world, rank)
    njobs = len(jobs)
    cands_per_job = (2 * R + 1) ** 2
    total_jobs = len(jobs) if args.virtual_world else ncx * ncy

    # pinned host frames in HM's sample type (Pel = int16), synthetic content
    f = luma_frames(W, H, 2)
    h_cur = torch.from_numpy(pad_plane(f[1], margin, margin)).pin_memory()
    h_ref = torch.from_numpy(pad_plane(f[0], margin, margin)).pin_memory()
    n_cur, n_ref = h_cur.numpy(), h_ref.numpy()
    band_h = 64 * (r1 - r0)
    n_cur_band = n_cur[margin + 64 * r0: margin + 64 * r1] if band_h else None
    pitch = (W + 2 * margin + 15) // 16 * 16
    rows = H + 2 * margin
    slice_rows = -(-rows // world)                      # reference rows each rank uploads when the plane is all-gathered
    s0, s1 = min(rank * slice_rows, rows), min((rank + 1) * slice_rows, rows)

    class Pipe:
        """One library context with its own stream, device planes (torch tensors, so NCCL can broadcast them; the library
        works on views) and page-locked result arrays."""

        def __init__(self):
            self.me = hm.MotionEstimator(local_rank, R)
            self.me.set_lambda_q16(LAMBDA_Q16)
            self.ext = torch.cuda.ExternalStream(self.me.stream_ptr, device=dev)
            self.t_cur = torch.zeros(rows * pitch + 64, dtype=torch.uint8, device=dev)
            self.t_ref = torch.zeros(slice_rows * world * pitch + 64, dtype=torch.uint8, device=dev)
            self.p_cur = self.me.wrap_plane(self.t_cur.data_ptr(), 1, pitch, W, H, margin, margin)
            self.p_ref = self.me.wrap_plane(self.t_ref.data_ptr(), 1, pitch, W, H, margin, margin)
            # this rank's band of the current frame as its own upload target (rows [64*r0, 64*r1), no vertical margin)
            self.p_cur_band = self.me.wrap_plane(self.t_cur.data_ptr() + (margin + 64 * r0) * pitch, 1, pitch, W, band_h, margin, 0) if band_h else None
            # this rank's horizontal slice of the (padded) reference plane as its own upload target
            self.p_ref_slice = self.me.wrap_plane(self.t_ref.data_ptr() + s0 * pitch, 1, pitch, W, s1 - s0, margin, 0) if s1 > s0 else None
            self.outs = [torch.zeros((max(njobs, 1), NPARTS), dtype=torch.int32).pin_memory().numpy().view(t)
                         for t in (np.int32, np.int32, np.uint32, np.uint32)]

        def upload_inputs(self, asynchronous=False):
            """The per-step host->device leg of the public API: reference picture (rank 0, then NCCL broadcast over NVLink),
            band of the current frame (every rank)."""
            if world > 1 and args.ref_dist == "allgather":
                # every rank pushes 1/N of the reference over its own PCIe link, then NCCL all-gathers the 8-bit slices in place
                if self.p_ref_slice is not None:
                    self.me.upload(self.p_ref_slice, n_ref[s0:s1], origin_x=margin, origin_y=0, asynchronous=asynchronous)
                with torch.cuda.stream(self.ext):
                    full = self.t_ref[:slice_rows * world * pitch]
                    dist.all_gather_into_tensor(full, full[rank * slice_rows * pitch:(rank + 1) * slice_rows * pitch])
            else:
                if rank == 0:
                    self.me.upload(self.p_ref, n_ref, asynchronous=asynchronous)
                if world > 1:
                    with torch.cuda.stream(self.ext):
                        dist.broadcast(self.t_ref, src=0)
            if band_h:
                self.me.upload(self.p_cur_band, n_cur_band, origin_x=margin, origin_y=0, asynchronous=asynchronous)

        def step_e2e(self, asynchronous):
            self.upload_inputs(asynchronous)
            if njobs:
                self.me.search_frame_async(self.p_cur, self.p_ref, jobs, R)
                self.me.fetch_results(njobs, self.outs, asynchronous=asynchronous)

        def bind(self):
            """Pre-marshalled calls of the pipelined e2e step (same C entry points as upload / search_frame_async / fetch_results):
            at N = 8 a step is 0.16 ms of kernels, so the Python argument handling of every call counts."""
            m = self.me
            self.b_sync = m.bind_sync()
            self.b_ref = None
            if world > 1 and args.ref_dist == "allgather":
                if self.p_ref_slice is not None:
                    self.b_ref = m.bind_upload(self.p_ref_slice, n_ref[s0:s1], origin_x=margin, origin_y=0)
            elif rank == 0:
                self.b_ref = m.bind_upload(self.p_ref, n_ref)
            self.b_cur = m.bind_upload(self.p_cur_band, n_cur_band, origin_x=margin, origin_y=0) if band_h else None
            self.b_search = m.bind_search(self.p_cur, self.p_ref, jobs, R) if njobs else None
            self.b_fetch = m.bind_fetch(njobs, self.outs) if njobs else None

        def step_e2e_bound(self):
            if self.b_ref is not None:
                self.b_ref()
            if world > 1:
                with torch.cuda.stream(self.ext):
                    if args.ref_dist == "allgather":
                        full = self.t_ref[:slice_rows * world * pitch]
                        dist.all_gather_into_tensor(full, full[rank * slice_rows * pitch:(rank + 1) * slice_rows * pitch])
                    else:
                        dist.broadcast(self.t_ref, src=0)
            if self.b_cur is not None:
                self.b_cur()
            if self.b_search is not None:
                self.b_search()
                self.b_fetch()

        def build_graphs(self):
            """The pipelined e2e step recorded as CUDA graphs (hmme_graph_*): at N = 8 a step is ~25 runtime calls for 0.16 ms of
            kernels.  One graph when the reference needs no collective, else {upload reference} -> NCCL -> {upload current, search, fetch}."""
            self.g_ref = self.g_main = None
            if world == 1:
                self.me.graph_begin()
                self.step_e2e(True)
                self.g_main = self.me.graph_end()
                return
            uploads_ref = (self.p_ref_slice is not None) if args.ref_dist == "allgather" else rank == 0
            if uploads_ref:
                self.me.graph_begin()
                if args.ref_dist == "allgather":
                    self.me.upload(self.p_ref_slice, n_ref[s0:s1], origin_x=margin, origin_y=0, asynchronous=True)
                else:
                    self.me.upload(self.p_ref, n_ref, asynchronous=True)
                self.g_ref = self.me.graph_end()
            if band_h and njobs:
                self.me.graph_begin()
                self.me.upload(self.p_cur_band, n_cur_band, origin_x=margin, origin_y=0, asynchronous=True)
                self.me.search_frame_async(self.p_cur, self.p_ref, jobs, R)
                self.me.fetch_results(njobs, self.outs, asynchronous=True)
                self.g_main = self.me.graph_end()

        def step_e2e_graph(self):
            if self.g_ref is not None:
                self.me.graph_launch(self.g_ref)
            if world > 1:
                with torch.cuda.stream(self.ext):
                    if args.ref_dist == "allgather":
                        full = self.t_ref[:slice_rows * world * pitch]
                        dist.all_gather_into_tensor(full, full[rank * slice_rows * pitch:(rank + 1) * slice_rows * pitch])
                    else:
                        dist.broadcast(self.t_ref, src=0)
            if self.g_main is not None:
                self.me.graph_launch(self.g_main)

        def step_e2e_frac(self):
            """Integer search + fractional refinement of all 593 partitions, host to host."""
            if not hasattr(self, "frac_out"):
                self.frac_out = torch.zeros((max(njobs, 1), NPARTS, 4), dtype=torch.int32).pin_memory().numpy()
            self.upload_inputs(True)
            if njobs:
                self.me.search_frame_async(self.p_cur, self.p_ref, jobs, R)
                self.me.fetch_results(njobs, self.outs, asynchronous=True)
                self.me.refine_frame(self.p_cur, self.p_ref, njobs, None, True, asynchronous=True, out=self.frac_out)

    pipes = [Pipe(), Pipe()]
    me, ext, p_cur, p_ref = pipes[0].me, pipes[0].ext, pipes[0].p_cur, pipes[0].p_ref
    upload_inputs = pipes[0].upload_inputs

    for pp in pipes:
        pp.upload_inputs()
        pp.me.sync()
    torch.cuda.synchronize()
    peak = me.measure_int_alu_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if njobs:
            me.search_frame_async(p_cur, p_ref, jobs, R)

    # ------------------------------------------------------------------ value: inputs resident in HBM
    # Resident inputs: NSETS copies of the (current, reference) plane pair at distinct addresses, together larger than twice
    # the 126 MB L2, cycled through step by step ("inputs larger than L2"; the kernel is compute bound, but the rule is kept).
    plane_stride = rows * pitch + 64
    nsets = max(2, -(-2 * 126 * (1 << 20) // (2 * rows * pitch)))
    t_sets = torch.zeros(2 * nsets * plane_stride, dtype=torch.uint8, device=dev)
    sets = []
    with torch.cuda.stream(ext):
        for k in range(nsets):
            oc, orf = (2 * k) * plane_stride, (2 * k + 1) * plane_stride
            t_sets[oc:oc + rows * pitch].copy_(pipes[0].t_cur[:rows * pitch])
            t_sets[orf:orf + rows * pitch].copy_(pipes[0].t_ref[:rows * pitch])
            sets.append((me.wrap_plane(t_sets.data_ptr() + oc, 1, pitch, W, H, margin, margin),
                         me.wrap_plane(t_sets.data_ptr() + orf, 1, pitch, W, H, margin, margin)))
    barrier()

    # (a) dominant-kernel duration and single-stream step time: one context, CUDA events per step on its stream
    for _ in range(args.warmup):
        step_resident()
    barrier()
    n_single = min(args.steps, 40)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_single)]
    kernel_ms = []
    with torch.cuda.stream(ext):
        for s in range(n_single):
            ev[s][0].record(ext)
            if njobs:
                me.search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
            ev[s][1].record(ext)
            if njobs:
                kernel_ms.append(me.last_kernel_ms())                  # CUDA events around the dominant kernel, same stream
    barrier()
    single_ms = sum(a.elapsed_time(b) for a, b in ev) / n_single

    # (b) the reported value: EXACTLY K steps, frames alternating over two contexts (two streams), nothing but the library's
    # kernels in the timed region; consecutive frames overlap at their wave tails.  CUDA events around the whole region.
    for s in range(max(args.warmup, 4)):
        if njobs:
            pipes[s & 1].me.search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = sum(pp.me.kernel_launches for pp in pipes)
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), [torch.cuda.Event(enable_timing=True) for _ in pipes]
    barrier()
    wall0 = time.perf_counter()
    ev_start.record(pipes[0].ext)
    for s in range(args.steps):
        if njobs:
            pipes[s & 1].me.search_frame_async(sets[s % nsets][0], sets[s % nsets][1], jobs, R)
    for pp, e_ in zip(pipes, ev_end):
        e_.record(pp.ext)
    barrier()
    wall1 = time.perf_counter()
    launches = sum(pp.me.kernel_launches for pp in pipes) - launches0
    total_ms = torch.tensor([max(ev_start.elapsed_time(e_) for e_ in ev_end), single_ms], dtype=torch.float64, device=dev)
    kern_ms = torch.tensor([float(np.mean(kernel_ms)) if kernel_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    total_ms, single_ms, kern_ms = float(total_ms[0].item()), float(total_ms[1].item()), float(kern_ms.item())
    clocks = sampler.summary() if sampler else None

    # ------------------------------------------------------------------ e2e: host buffers through the public API
    # (a) serial: every step waits for its own results before the next upload starts (a low-delay encoder's dependency);
    # (b) pipelined (the reported e2e): frames alternate between two contexts, so the H2D/D2H copies of one frame overlap
    #     the kernels of the other -- every step still uploads both int16 planes and downloads its four result arrays.
    for _ in range(2):
        pipes[0].step_e2e(False)
        pipes[1].step_e2e(False)
    barrier()
    e0 = time.perf_counter()
    for s in range(args.steps):
        pipes[0].step_e2e(False)
    barrier()
    serial_ms = (time.perf_counter() - e0) * 1e3
    use_graphs = args.graphs
    if use_graphs:
        for pp in pipes:
            pp.me.sync()
            pp.build_graphs()
        for pp in pipes:                               # one replay each before the clock starts
            pp.step_e2e_graph()
            pp.me.sync()
        barrier()
    for pp in pipes:
        pp.bind()
    barrier()
    e0 = time.perf_counter()
    for s in range(args.steps):
        pp = pipes[s & 1]
        pp.b_sync()                                    # this context's previous frame (two steps ago) is complete
        if use_graphs:
            pp.step_e2e_graph()
        else:
            pp.step_e2e_bound()
    for pp in pipes:
        pp.me.sync()
    barrier()
    e1 = time.perf_counter()
    e2e_ms = torch.tensor([(e1 - e0) * 1e3, serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms, serial_ms = float(e2e_ms[0].item()), float(e2e_ms[1].item())
    ref_bytes = n_ref[s0:s1].nbytes if (world > 1 and args.ref_dist == "allgather") else (n_ref.nbytes if rank == 0 else 0)
    h2d = ref_bytes + (n_cur_band.nbytes if band_h else 0) + jobs.nbytes
    d2h = 4 * njobs * NPARTS * 4
    io = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(io, op=dist.ReduceOp.SUM)

    # ------------------------------------------------------------------ next row: fractional-pel refinement
    frac = None
    if not args.virtual_world:
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
        ev_a.record(ext)
        for s in range(nfr):                                # search + refinement per frame, one context: the two kernels of a frame
            pc_, pr_ = sets[s % nsets]                      # depend on each other and both fill the GPU, so there is nothing to overlap
            if njobs:                                       # (two contexts with deep queues showed erratic CTA interleaving: 2.5-5 ms)
                me.search_frame_async(pc_, pr_, jobs, R)
                me.refine_frame(pc_, pr_, njobs, None, True, asynchronous=True)
        ev_b.record(ext)
        barrier()
        both_ms = ev_a.elapsed_time(ev_b) / nfr
        for pp in pipes:
            pp.step_e2e_frac(); pp.me.sync()
        e0 = time.perf_counter()
        for s in range(nfr):
            pp = pipes[s & 1]
            pp.me.sync()
            pp.step_e2e_frac()
        for pp in pipes:
            pp.me.sync()
        barrier()
        e2e_frac_ms = (time.perf_counter() - e0) * 1e3 / nfr
        ft = torch.tensor([float(np.mean(fk)), float(np.mean(fk_sad)), both_ms, e2e_frac_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fk, fk_sad, both_ms, e2e_frac_ms = [float(ft[0].item())], [float(ft[1].item())], float(ft[2].item()), float(ft[3].item())
        pu_px = total_jobs * 24 * 4096                      # sum of the 593 partition areas = 24 CTU areas
        # algorithmic operations of the reference's own scheme per CTU (all 593 partitions, half-pel winner at the centre): filter MACs over
        # the plane sizes of xExtDIFUpSamplingH/Q + 8 (8x8 Hadamard) or 6 (4x4) operations per pixel and candidate; formula in DESIGN.md 3.4
        frac_ops = FRAC_OPS_PER_CTU * njobs
        frac_peak = 2.0 * peak["lane_ops_per_s"]            # both integer pipes (ALU + FMA-heavy/IMAD), each at the measured 64 lanes/clk/SM
        frac = {"scope": "fractional-pel refinement (xPatternSearchFracDIF: 9 half-pel + 9 quarter-pel candidates, 8-tap interpolation, Hadamard cost) "
                         "of all 593 partitions of every CTU, from the integer winners left on the device",
                "kernel": "me_frac_kernel", "pus_per_frame": total_jobs * NPARTS, "kernel_ms": float(np.mean(fk)), "kernel_ms_sad": float(np.mean(fk_sad)),
                "pu_refinements_per_s": total_jobs * NPARTS / (np.mean(fk) * 1e-3), "pu_pixels_per_s": pu_px / (np.mean(fk) * 1e-3),
                "search_plus_refine_ms_per_frame": both_ms, "search_plus_refine_frames_per_s": 1e3 / both_ms,
                "e2e_ms_per_frame": e2e_frac_ms, "e2e_frames_per_s": 1e3 / e2e_frac_ms,
                "e2e_d2h_bytes_per_step": 4 * total_jobs * NPARTS * 4 + total_jobs * NPARTS * 16, "steps": nfr,
                "roofline": {"bound": "int_issue", "achieved": frac_ops / (np.mean(fk) * 1e-3) / 1e12 if np.mean(fk) > 0 else 0.0, "peak": frac_peak / 1e12,
                             "unit": "T int-op/s", "frac": (frac_ops / (np.mean(fk) * 1e-3)) / frac_peak if np.mean(fk) > 0 and frac_peak else None,
                             "ops_per_ctu": FRAC_OPS_PER_CTU, "peak_source": "2 x the live-measured integer-ALU issue rate (ALU and FMA-heavy pipes issue concurrently)"},
                "timer": "kernel_ms: CUDA events around me_frac_kernel on its stream; search_plus_refine: CUDA events around K frames on one context/stream, "
                         "resident inputs; e2e: host wall clock, uploads + search + refine + all result arrays fetched, two contexts"}

    # ------------------------------------------------------------------ row f3: motion-compensated distortion at quarter-pel MVs
    mc = None
    if world == 1 and njobs and not args.virtual_world:
        rng = np.random.default_rng(7)
        rects = me.lib.partition_table()
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
              "kernel": "me_mc_cost_kernel", "pus": njobs * NPARTS, "kernel_ms_sad": km["sad"], "kernel_ms_hadamard": km["hadamard"], "kernel_ms_bi_hadamard": km["bi_hadamard"],
              "pu_pixels_per_s_sad": njobs * 24 * 4096 / (km["sad"] * 1e-3), "timer": "CUDA events around the kernel on its stream, resident planes"}

    if rank == 0:
        total_cands = total_jobs * cands_per_job
        ms_per_step = total_ms / args.steps
        value = total_cands * NPARTS / (ms_per_step * 1e-3)
        e2e_value = total_cands * NPARTS / (e2e_ms / args.steps * 1e-3)
        # dominant kernel roofline: algorithmic integer lane-ops of THIS rank's launch / its CUDA-event duration
        cands_rank = njobs * cands_per_job
        achieved = cands_rank * INT_OPS_PER_CAND / (kern_ms * 1e-3) if kern_ms > 0 else 0.0
        alg_bytes = (W + 2 * margin) * (H + 2 * margin) + W * H + total_jobs * NPARTS * 16
        out = {
            "metric": "me_block_sad_evaluations_per_s", "value": value, "unit": "block-SAD evaluations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "frames_per_s": 1e3 / ms_per_step,
            "ctu_candidates_per_s": total_cands / (ms_per_step * 1e-3),
            "config": {"workload": "%s: %dx%d luma, 64x64 CTU, integer-pel full search +-%d, 1 reference picture, %d CTU jobs x %d candidates x 593 partitions"
                                   % (args.workload, W, H, R, total_jobs, cands_per_job),
                       "sharding": "CTU-row bands (cut at CTU granularity) over %d GPU(s); reference plane: %s" % (world, "single GPU" if world == 1 else ("each rank uploads 1/N, NCCL all-gather over NVLink" if args.ref_dist == "allgather" else "rank 0 uploads, NCCL broadcast over NVLink")),
                       "lambda_q16": LAMBDA_Q16,
                       "l2": "inputs larger than L2: %d resident (current, reference) plane pairs at distinct addresses (%.0f MiB), cycled step by step" % (nsets, 2 * nsets * plane_stride / 2**20),
                       "timer": "CUDA events around the whole K-step region, frames alternating over two library contexts/streams, max over ranks; "
                                "single_stream_ms_per_step = one context, events per step"},
            "single_stream_ms_per_step": single_ms,
            "clocks": clocks,
            "gpu_launches": int(launches),
            "wall_ms_timed_region": (wall1 - wall0) * 1e3,
            "e2e": {"value": e2e_value, "unit": "block-SAD evaluations/s", "h2d_bytes_per_step": int(io[0].item()),
                    "d2h_bytes_per_step": int(io[1].item()), "frames_per_s": 1e3 / (e2e_ms / args.steps),
                    "ms_per_step": e2e_ms / args.steps,
                    "serial_ms_per_step": serial_ms / args.steps, "serial_frames_per_s": 1e3 / (serial_ms / args.steps),
                    "timer": "host wall clock around K x {upload both int16 planes from pinned memory (+ NCCL broadcast), search, fetch four result arrays}, "
                             "frames alternating over two contexts/streams so copies overlap kernels" + ("; each context's step is recorded once as CUDA graph(s) and replayed (hmme_graph_*)" if use_graphs else "") +
                             "; serial_* = one context, call by call, each step waits for its results; max over ranks"},
            "roofline": {"bound": "int_alu", "kernel": "me_u8_tile_kernel", "achieved": achieved / 1e12, "peak": peak["lane_ops_per_s"] / 1e12,
                         "unit": "T int-lane-op/s", "frac": achieved / peak["lane_ops_per_s"] if peak["lane_ops_per_s"] else None,
                         "traffic": _ncu_traffic(args.workload) if world == 1 else None,
                         "ops_per_ctu_candidate": INT_OPS_PER_CAND, "kernel_ms": kern_ms,
                         "frac_at_step_rate": (cands_rank * INT_OPS_PER_CAND / (ms_per_step * 1e-3)) / peak["lane_ops_per_s"] if peak["lane_ops_per_s"] else None,
                         "pixel_abs_diffs_per_s": cands_rank * PX_PER_CAND / (kern_ms * 1e-3) if kern_ms > 0 else 0.0,
                         "peak_source": "measured live: VABSDIFF4.U8.ACC issue rate, %.1f lanes/clk/SM x %d SMs at %.0f MHz"
                                        % (peak["lanes_per_clk_sm"], torch.cuda.get_device_properties(dev).multi_processor_count, peak["sm_mhz"]),
                         "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0,
                                 "peak_gbs": _measured_hbm()}},
        }
        if frac:
            out["frac_refine"] = frac
        if mc:
            out["mc_cost"] = mc
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], _ = cpu_baseline_entry(R, 30.0)
            if frac and os.path.exists(CPUME_BIN):
                fr = reference_cpu_frac()
                frac["cpu_reference"] = fr
                frac["gpu_over_one_core"] = frac["pu_pixels_per_s"] / fr["pu_pixels_per_s"]
            if os.path.exists(CPUME_BIN):       # the reference's default (fast) integer search, for context: TZ evaluates ~600x fewer candidates
                v, me_s, calls, wall = reference_cpu_me(R, (416, 240), 1, fast_search=1)
                out["cpu_tz"] = {"value": v, "unit": "block-SAD evaluations/s", "cores": 1, "kind": "reference",
                                 "sample": "same encoder, --FastSearch=1 (xTZSearch): %d DistFunc calls in %.3f s of integer ME for the P frame of a "
                                           "416x240 clip (18 full CTUs) on one core" % (calls, me_s)}
            threads = os.cpu_count() or 1
            v, dt, n = cpu_oracle_throughput(W, H, R, margin, max(8 * threads, 64), threads)
            out["cpu_port"] = {"value": v, "unit": "block-SAD evaluations/s", "cores": threads, "kind": "port",
                               "sample": "GPU-ME semantics on the CPU (oracle/hmme_oracle.c): %d of %d CTU jobs of the same frame, %.1f s wall on %d threads" % (n, total_jobs, dt, threads)}
        print(json.dumps(out))
    for pp in pipes:
        pp.me.close()
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return t["traffic_bytes_per_launch"] if t["workload"].startswith(workload) else None
    except Exception:
        return None


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0   # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
