"""N > 1 host logic on the CPU: two gloo ranks shard a frame into CTU-row bands (hm.band_jobs), rank 0 broadcasts the
reference plane, each rank searches its band (with the oracle standing in for the device here -- there is no GPU in this
container) and the gathered result must equal the single-rank result.  Mirrors what bench.py does over NCCL."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import hm
    from oracle.pyoracle import Oracle
    from synth import luma_frames, pad_plane
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, R, M, lam = 256, 320, 5, 24, 460000          # 4 x 5 CTUs: bands of 3 and 2 rows
    f = luma_frames(W, H, 2, seed=9)
    cur = pad_plane(f[1], M, M)
    ref = torch.from_numpy(pad_plane(f[0], M, M)) if rank == 0 else torch.zeros((H + 2 * M, W + 2 * M), dtype=torch.int16)
    dist.broadcast(ref.view(torch.uint8), src=0)        # reference-picture distribution (bytes, like the u8 plane over NCCL)
    jobs, (r0, r1) = hm.band_jobs(W, H, R, world, rank)
    c0, c1 = hm.band_ctus((W // 64) * (H // 64), world, rank)
    assert len(jobs) == c1 - c0 and (r0, r1) == (c0 // (W // 64), (c1 - 1) // (W // 64) + 1)
    lib = hm.HmmeLib.get()                              # the library's own band arithmetic (what hmme_group_* uses) agrees with the helpers
    assert lib.band_split((W // 64) * (H // 64), world, rank) == (c0, c1 - c0)
    y0, y1 = hm.band_reference_rows(r0, r1, R, -R, -R)
    cr, rr = lib.band_extent(jobs, R)
    assert (rr[1], rr[3]) == (y0, y1) and (cr[1], cr[3]) == (64 * r0, 64 * r1)
    assert y0 >= -M and y1 <= H + M                     # the halo stays inside the padded plane
    ref_np = ref.numpy().copy()                         # band + halo distribution: everything outside the rectangle may be garbage
    keep = np.zeros_like(ref_np, dtype=bool)
    keep[M + rr[1]:M + rr[3], M + rr[0]:M + rr[2]] = True
    ref_np[~keep] = -9999
    ref = torch.from_numpy(ref_np)
    res = Oracle().search_frame(cur, (M, M), ref.numpy(), (M, M), jobs, R, lam)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), X=res[0], Y=res[1], S=res[2], C=res[3], r=np.array([r0, r1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_sharding_matches_single_rank(tmp_path, oracle):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import hm
    from synth import frame_jobs, luma_frames, pad_plane
    world = 2
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 2000, str(tmp_path)), nprocs=world, join=True)
    parts = []
    rows = []
    for r in range(world):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        parts.append((z["X"], z["Y"], z["S"], z["C"]))
        rows.append(tuple(z["r"]))
    assert rows == [(0, 3), (2, 5)]          # 20 CTUs -> 10 + 10: the middle row is shared
    got = hm.merge_bands(parts)
    W, H, R, M, lam = 256, 320, 5, 24, 460000
    f = luma_frames(W, H, 2, seed=9)
    want = oracle.search_frame(pad_plane(f[1], M, M), (M, M), pad_plane(f[0], M, M), (M, M), frame_jobs(W, H, R), R, lam, nthreads=4)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_band_ctus_cover_exactly_and_balance():
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import hm
    for n in (1, 480, 1980, 18):
        for world in (1, 2, 4, 8):
            bands = [hm.band_ctus(n, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == n and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1
    jobs, rows = hm.band_jobs(3840, 2160, 128, 8, 0)
    assert len(jobs) == 248 and rows == (0, 5)


def test_band_rows_cover_exactly():
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import hm
    for n in (1, 16, 17, 33):
        for world in (1, 2, 4, 8):
            bands = [hm.band_rows(n, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1


def test_library_band_split_and_extent_on_the_baseline_shapes():
    """hmme_band_split / hmme_band_extent (pure host code of the C ABI, used by hmme_group_*): exact cover, balance, and the band + halo
    rectangles of BASELINE configs 2 and 4 at 8 GPUs."""
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import hm
    from synth import frame_jobs
    lib = hm.HmmeLib.get()
    for n in (1, 5, 18, 480, 1980):
        for world in (1, 2, 4, 8):
            bands = [lib.band_split(n, world, r) for r in range(world)]
            assert [b for b in bands] == [(hm.band_ctus(n, world, r)[0], hm.band_ctus(n, world, r)[1] - hm.band_ctus(n, world, r)[0]) for r in range(world)]
            assert sum(c for _, c in bands) == n
    jobs = frame_jobs(1920, 1080, 64)
    f, c = lib.band_split(len(jobs), 8, 7)
    cr, rr = lib.band_extent(jobs[f:f + c], 64)
    assert (f, c) == (420, 60) and cr == (0, 896, 1920, 1024) and rr == (-64, 832, 1984, 1088)
    rows = rr[3] - rr[1]
    assert rows / (1080 + 160) < 0.27                    # each of 8 GPUs receives about a quarter of the padded plane
    jobs = frame_jobs(3840, 2160, 128)
    f, c = lib.band_split(len(jobs), 8, 0)
    cr, rr = lib.band_extent(jobs[f:f + c], 128)
    assert c == 248 and cr == (0, 0, 3840, 320) and rr == (-128, -128, 3968, 448)
    import pytest
    with pytest.raises(hm.HmmeError):
        lib.band_split(10, 4, 4)
