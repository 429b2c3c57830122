"""CTU-row band sharding of a frame over the GPUs of one box.

The path has no data-path exchange step: every (CTU, reference) job is independent once the reference plane is on the
device, so ranks take contiguous bands of CTU rows (cut at CTU granularity), rank 0's reference upload is NCCL-broadcast, and results come back
per rank.  (The reference itself has no multi-device code at all: one queue on one device, TEncOpenCL.cpp:185.)
"""
import numpy as np


def band_rows(n_ctu_rows, world, rank):
    """Contiguous CTU-row band [r0, r1) of `rank`; band sizes differ by at most one row (1080p: 16 rows -> 2 per GPU at 8)."""
    base, extra = divmod(n_ctu_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def band_ctus(n_ctus, world, rank):
    """Contiguous raster range [c0, c1) of CTUs for `rank`: a CTU-row band whose first and last row may be partial, so
    that every rank gets the same number of jobs (+-1) even when the row count does not divide (4K: 33 rows over 8 GPUs
    would otherwise leave one rank with 5 rows against 4)."""
    base, extra = divmod(n_ctus, world)
    c0 = rank * base + min(rank, extra)
    return c0, c0 + base + (1 if rank < extra else 0)


def band_jobs(width, height, search_range, world, rank, pred=(0, 0)):
    """Jobs {ctuX, ctuY, ltx, lty} of this rank's band: one per FULL 64x64 CTU (partial boundary CTUs never run the
    depth-0 search, TEncCu.cpp:424-425), raster order, window centred on `pred`.  Returns (jobs, (r0, r1)) with
    [r0, r1) the CTU rows the band touches (the rows of the current frame this rank has to upload)."""
    nx, ny = width // 64, height // 64
    c0, c1 = band_ctus(nx * ny, world, rank)
    jobs = [[(c % nx) * 64, (c // nx) * 64, pred[0] - search_range, pred[1] - search_range] for c in range(c0, c1)]
    rows = (c0 // nx, (c1 - 1) // nx + 1) if c1 > c0 else (0, 0)
    return np.asarray(jobs, np.int32).reshape(-1, 4), rows


def band_reference_rows(r0, r1, search_range, lty_min, lty_max):
    """Picture rows [y0, y1) of the reference plane a band reads: its own rows plus the halo of the search window
    (a CTU at row y reads rows [y + lty, y + lty + 2R + 63]).  Only needed when a rank uploads the band instead of receiving the whole plane."""
    return 64 * r0 + lty_min, 64 * (r1 - 1) + lty_max + 2 * search_range + 63 + 1


def merge_bands(parts):
    """Concatenate per-rank result tuples (X, Y, sad, cost), in rank order, back into frame (raster) order."""
    return tuple(np.concatenate([p[k] for p in parts if len(p[k])], axis=0) for k in range(4))
