"""CTU-row band sharding of a frame over the GPUs of one box.

The path has no data-path exchange step: every (CTU, reference) job is independent once the reference plane is on the
device, so ranks take contiguous bands of CTU rows, rank 0's reference upload is NCCL-broadcast, and results come back
per rank.  (The reference itself has no multi-device code at all: one queue on one device, TEncOpenCL.cpp:185.)
"""
import numpy as np


def band_rows(n_ctu_rows, world, rank):
    """Contiguous CTU-row band [r0, r1) of `rank`; band sizes differ by at most one row (1080p: 16 rows -> 2 per GPU at 8)."""
    base, extra = divmod(n_ctu_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def band_jobs(width, height, search_range, world, rank, pred=(0, 0)):
    """Jobs {ctuX, ctuY, ltx, lty} of this rank's band: one per FULL 64x64 CTU (partial boundary CTUs never run the
    depth-0 search, TEncCu.cpp:424-425), raster order, window centred on `pred`."""
    nx, ny = width // 64, height // 64
    r0, r1 = band_rows(ny, world, rank)
    jobs = [[cx * 64, cy * 64, pred[0] - search_range, pred[1] - search_range] for cy in range(r0, r1) for cx in range(nx)]
    return np.asarray(jobs, np.int32).reshape(-1, 4), (r0, r1)


def band_reference_rows(r0, r1, search_range, lty_min, lty_max):
    """Picture rows [y0, y1) of the reference plane a band reads: its own rows plus the halo of the search window
    (a CTU at row y reads rows [y + lty, y + lty + 2R + 63]).  Only needed when a rank uploads the band instead of receiving the whole plane."""
    return 64 * r0 + lty_min, 64 * (r1 - 1) + lty_max + 2 * search_range + 63 + 1


def merge_bands(parts):
    """Concatenate per-rank result tuples (X, Y, sad, cost), in rank order, back into frame (raster) order."""
    return tuple(np.concatenate([p[k] for p in parts if len(p[k])], axis=0) for k in range(4))
