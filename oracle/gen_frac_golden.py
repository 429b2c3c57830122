#!/usr/bin/env python
"""Golden records for the fractional-pel refinement FROM THE REFERENCE ITSELF.

The instrumented reference encoder (oracle/_ref/TAppEncoder_cpume, see patch_cpume.py) appends one binary record per sampled
TEncSearch::xPatternSearchFracDIF call (TEncSearch.cpp:4294-4331): block size, bi-prediction flag, Hadamard flag, integer MV,
predictor, lambda (TComRdCost::m_uiCost), the current block, the reference patch with a 4-sample apron, and what the function
returned (half-pel winner, quarter-pel winner, cost), plus the cost of each of the 9 + 9 candidates it evaluated.  They are stored in tests/golden/frac_records.npz; tests check the
oracle (and the CUDA path) against them.  Needs /root/reference and `make -C oracle encoders`.  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_encoder_golden import REFDIR, write_yuv  # noqa: E402

RUNS = [  # (W, H, frames, cfg, extra args, per-class cap, sampling stride, sub-pel motion clip?)
    (416, 240, 5, "encoder_randomaccess_main.cfg", ["--SearchRange=16"], 2, 97, False),
    (416, 240, 3, "encoder_lowdelay_P_main.cfg", ["--SearchRange=32", "-q", "27"], 1, 61, False),
    (192, 128, 3, "encoder_randomaccess_main.cfg", ["--SearchRange=16", "--HadamardME=0"], 1, 41, False),
    (256, 192, 5, "encoder_randomaccess_main.cfg", ["--SearchRange=16", "-q", "24"], 3, 53, True),
    (256, 192, 3, "encoder_lowdelay_P_main.cfg", ["--SearchRange=16", "--HadamardME=0", "-q", "27"], 1, 37, True),
]
HDR = ["magic", "w", "h", "bi", "had", "mvx", "mvy", "predx", "predy", "lambda", "halfx", "halfy", "qterx", "qtery", "cost", "pad"]


def write_subpel_yuv(path, W, H, F, seed=77):
    """Smooth texture moving by (0.75, -0.5) samples per frame (bilinear resampling of a 4x finer canvas), so that the
    half- and quarter-pel stages have non-trivial winners."""
    rng = np.random.default_rng(seed)
    fine = rng.integers(0, 256, size=((H + 32) * 4, (W + 32) * 4)).astype(np.float64)
    for _ in range(3):                                   # separable box blurs -> band-limited texture
        fine = (np.roll(fine, 3, 0) + np.roll(fine, -3, 0) + np.roll(fine, 6, 0) + np.roll(fine, -6, 0) + fine) / 5
        fine = (np.roll(fine, 3, 1) + np.roll(fine, -3, 1) + np.roll(fine, 6, 1) + np.roll(fine, -6, 1) + fine) / 5
    fine = np.clip((fine - fine.mean()) * 6 + 128, 0, 255)
    with open(path, "wb") as f:
        for t in range(F):
            ox, oy = 32 + 3 * t, 48 - 2 * t              # quarter-sample units on the fine grid
            y = fine[oy:oy + 4 * H:4, ox:ox + 4 * W:4]
            f.write(np.round(y).astype(np.uint8).tobytes())
            f.write(np.full((H // 2) * (W // 2) * 2, 128, np.uint8).tobytes())


def parse(path):
    raw = np.fromfile(path, np.int16)
    recs, pos = [], 0
    while pos < raw.size:
        hdr = raw[pos:pos + 32].view(np.int32).copy()
        assert hdr[0] == 0x46524143, hex(int(hdr[0]))
        w, h = int(hdr[1]), int(hdr[2])
        pos += 32
        costs = raw[pos:pos + 36].view(np.uint32).copy()        # the 9 + 9 candidate costs in table order
        pos += 36
        cur = raw[pos:pos + w * h].reshape(h, w).copy()
        pos += w * h
        patch = raw[pos:pos + (w + 8) * (h + 8)].reshape(h + 8, w + 8).copy()
        pos += (w + 8) * (h + 8)
        recs.append((hdr, cur, patch, costs))
    return recs


def check_with_oracle(recs):
    """Oracle vs. the reference's own outputs, record by record (the pin)."""
    from oracle.pyoracle import Oracle
    O = Oracle()
    bad = 0
    for hdr, cur, patch, costs in recs:
        w, h = int(hdr[1]), int(hdr[2])
        # the patch is its own little reference plane whose sample (0,0) is the PU origin displaced by the integer MV
        pu = np.array([[0, 0, w, h, 0, 0, int(hdr[7]) - 4 * int(hdr[5]), int(hdr[8]) - 4 * int(hdr[6])]], np.int32)
        r = O.refine_frac(np.ascontiguousarray(cur), (0, 0), np.ascontiguousarray(patch), (4, 4), pu, int(np.uint32(hdr[9])), bool(hdr[4]))
        got = (r["half"][0, 0], r["half"][0, 1], r["qter"][0, 0], r["qter"][0, 1], int(r["cost"][0]))
        want = (hdr[10], hdr[11], hdr[12], hdr[13], int(np.uint32(hdr[14])))
        if tuple(int(v) for v in got) != tuple(int(v) for v in want) or not np.array_equal(r["cand"][0], costs):
            bad += 1
            if bad < 10:
                print("MISMATCH", dict(zip(HDR, hdr.tolist())), "oracle", got)
    return bad


def main():
    binary = os.path.join(REFDIR, "TAppEncoder_cpume")
    recs = []
    with tempfile.TemporaryDirectory() as d:
        for W, H, F, cfg, extra, cap, stride, subpel in RUNS:
            yuv, log = os.path.join(d, "c.yuv"), os.path.join(d, "frac.bin")
            (write_subpel_yuv if subpel else write_yuv)(yuv, W, H, F)
            if os.path.exists(log):
                os.remove(log)
            env = dict(os.environ, HMME_LOG_FRAC=log, HMME_LOG_FRAC_CAP=str(cap), HMME_LOG_FRAC_STRIDE=str(stride))
            r = subprocess.run([binary, "-c", os.path.join(REFDIR, "cfg", cfg), "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(F),
                                "-q", "32", "-b", os.path.join(d, "o.hevc"), "-o", os.path.join(d, "rec.yuv"), "--OpenCL=0"] + extra,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
            assert r.returncode == 0, r.stdout[-2000:]
            got = parse(log)
            print(cfg, W, H, extra, "->", len(got), "records")
            recs += got
    bad = check_with_oracle(recs)
    print("oracle vs reference records: %d mismatches of %d" % (bad, len(recs)))
    assert bad == 0
    hdr = np.stack([r[0] for r in recs]).astype(np.int32)
    cur = np.concatenate([r[1].ravel() for r in recs]).astype(np.int16)
    patch = np.concatenate([r[2].ravel() for r in recs]).astype(np.int16)
    out = os.path.join(ROOT, "tests", "golden", "frac_records.npz")
    cand = np.stack([r[3] for r in recs]).astype(np.uint32)
    np.savez_compressed(out, columns=np.array(HDR), hdr=hdr, cur=cur, patch=patch, cand=cand)
    print("wrote", out, os.path.getsize(out), "bytes;", len(recs), "records; classes",
          sorted(set((int(h[1]), int(h[2]), int(h[3]), int(h[4])) for h in hdr)))


if __name__ == "__main__":
    main()
