"""The multi-GPU entry points of the C ABI (hmme_group_*, include/hmme_b200.h) on real hardware: band split, band + halo and
NCCL-broadcast reference distribution, results into one host table -- every CTU compared with the CPU oracle.
One GPU is enough for most of it (a rank of a larger world only ever touches its own band); the two-GPU cases run when
two devices are visible (the driver's box has 8)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from _pkg import hm
from synth import frame_jobs, luma_frames, pad_plane
from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    return hm.HmmeLib.get().device_count()


def frame(W, H, M, seed):
    f = luma_frames(W, H, 2, seed=seed)
    return pad_plane(f[1], M, M), pad_plane(f[0], M, M)


@pytest.mark.parametrize("dtype", [np.int16, np.uint8])
def test_group_of_one_equals_oracle(oracle, dtype):
    W, H, R, M, lam = 448, 320, 12, 32, 460000
    cur, ref = frame(W, H, M, 3)
    jobs = frame_jobs(W, H, R)
    jobs[::3, 2] += 4
    g = hm.Group(devices=[0], max_search_range=64)
    g.set_lambda_q16(lam)
    g.configure(W, H, M, M, g.BAND_HALO)
    got = g.search_frame(cur.astype(dtype), (M, M), ref.astype(dtype), (M, M), jobs, R)
    assert_same(got, oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8), "group of one")
    g.configure(W, H, M, M, g.BROADCAST)                 # a world of one needs no collective: same call, same result
    got = g.search_frame(cur.astype(dtype), (M, M), ref.astype(dtype), (M, M), jobs, R)
    assert_same(got, oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8), "group of one, broadcast mode")
    g.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_rank_of_a_world_fills_its_band_of_one_table(oracle, world):
    """One process per rank (hmme_group_create_rank), here one after the other on GPU 0: every rank passes the WHOLE frame and the
    same host tables; it uploads only its band + halo rectangles into zeroed device planes and writes only its band's rows.
    After the last rank the table must equal the oracle's whole-frame result."""
    W, H, R, M, lam = 448, 320, 12, 28, 262144            # 7 x 5 CTUs: bands cut mid-row
    cur, ref = frame(W, H, M, 10 + world)
    jobs = frame_jobs(W, H, R)
    jobs[1::2, 3] -= 5
    outs = hm.MotionEstimator._outs(len(jobs))
    for o in outs:
        o[:] = 0x55
    covered = 0
    for rank in range(world):
        g = hm.Group(device=0, rank=rank, world=world, unique_id=None, max_search_range=16)
        g.set_lambda_q16(lam)
        g.configure(W, H, M, M, g.BAND_HALO)
        first, n = g.band(len(jobs))
        assert (first, n) == g.lib.band_split(len(jobs), world, rank)
        before = [o.copy() for o in outs]
        g.search_frame_async(rank & 1, cur, (M, M), ref, (M, M), jobs, R, outs)
        g.sync()
        for o, b in zip(outs, before):                   # nothing outside the band's rows was written
            assert np.array_equal(o[:first], b[:first]) and np.array_equal(o[first + n:], b[first + n:])
        covered += n
        g.close()
    assert covered == len(jobs)
    assert_same(outs, oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8), f"world of {world}")


def test_group_errors_are_loud():
    W, H, M = 128, 128, 16
    cur, ref = frame(W, H, M, 1)
    g = hm.Group(devices=[0], max_search_range=16)
    outs = hm.MotionEstimator._outs(4)
    with pytest.raises(hm.HmmeError) as e:               # not configured
        g.search_frame_async(0, cur, (M, M), ref, (M, M), frame_jobs(W, H, 8), 8, outs)
    assert e.value.code == -1
    g.configure(W, H, M, M, g.BAND_HALO)
    with pytest.raises(hm.HmmeError) as e:               # window leaves the padded picture
        g.search_frame_async(0, cur, (M, M), ref, (M, M), frame_jobs(W, H, 16, pred=(9, 0)), 16, outs)
    assert e.value.code == -6
    with pytest.raises(hm.HmmeError) as e:               # beyond the range the group was created for
        g.search_frame_async(0, cur, (M, M), ref, (M, M), frame_jobs(W, H, 8), 32, outs)
    assert e.value.code == -4
    bad = ref.copy(); bad[M + 5, M + 5] = 999             # not 8-bit video: reported by the sync, the group stays usable
    g.search_frame_async(0, cur, (M, M), bad, (M, M), frame_jobs(W, H, 8), 8, outs)
    with pytest.raises(hm.HmmeError) as e:
        g.sync()
    assert e.value.code == -5
    g.search_frame(cur, (M, M), ref, (M, M), frame_jobs(W, H, 8), 8)
    g.close()
    r = hm.Group(device=0, rank=1, world=2, unique_id=None, max_search_range=16)
    with pytest.raises(hm.HmmeError):                    # a broadcast needs the communicator id
        r.configure(W, H, M, M, r.BROADCAST)
    r.close()
    with pytest.raises(hm.HmmeError):
        hm.Group(devices=[0, 0], max_search_range=16)


@pytest.mark.parametrize("mode", ["band_halo", "broadcast"])
def test_two_gpus_one_process(oracle, mode):
    """hmme_group_create over two real GPUs: one enqueue thread per GPU, ncclCommInitAll + ncclBroadcast inside the library."""
    if ngpus() < 2:
        pytest.skip("needs two GPUs")
    W, H, R, M, lam = 1920, 1080, 64, 80, 460000
    cur, ref = frame(W, H, M, 1234)
    jobs = frame_jobs(W, H, R)
    g = hm.Group(devices=[0, 1], max_search_range=64)
    g.set_lambda_q16(lam)
    g.configure(W, H, M, M, g.BROADCAST if mode == "broadcast" else g.BAND_HALO)
    n = g.SLOTS
    outs = [hm.MotionEstimator._outs(len(jobs)) for _ in range(n)]
    for s in range(2 * n):                               # every slot, twice: pipelined frames
        g.search_frame_async(s % n, cur, (M, M), ref, (M, M), jobs, R, outs[s % n])
        if s >= n - 1:
            g.sync((s - (n - 1)) % n)
    g.sync()
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=os.cpu_count() or 8)
    for o in outs:
        assert_same(o, want, f"two GPUs, {mode}")
    g.close()


def test_torchrun_two_ranks_bench_verifies_every_ctu():
    """The driver's own N > 1 launch of bench.py: two ranks over NCCL, every CTU of both bands verified inside bench.py."""
    if ngpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29671",
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "4", "--warmup", "3"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["n_gpus"] == 2 and d["verified"]["ctus"] == 480 and d["verified"]["mismatches"] == 0
    assert d["verified"]["ctu_result_sets_compared"] == (d["e2e"]["slots"] + 1) * 480
    assert d["e2e"]["band_halo_u8"]["value"] > 0 and d["e2e"]["broadcast_s16"]["value"] > 0
