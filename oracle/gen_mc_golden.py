#!/usr/bin/env python
"""Golden records for the motion-compensated template cost FROM THE REFERENCE ITSELF.

The instrumented reference encoder (oracle/_ref/TAppEncoder_cpume, see patch_cpume.py) appends one binary record per sampled
TEncSearch::xGetTemplateCost call (TEncSearch.cpp:3634-3674): block size, the clipped AMVP candidate MV (quarter pel), the SAD it
computed between the original block and xPredInterBlk's prediction, the block and the reference patch around the MV's integer
part; and one per sampled xGetInterPredictionError call (:2814-2836: merge candidates and the motion-estimation result, uni- and
bi-directional, Hadamard).  Stored in tests/golden/mc_records.npz; tests check the oracle and the CUDA path against them.
Needs /root/reference and `make -C oracle encoders`.  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_encoder_golden import REFDIR, write_yuv  # noqa: E402
from oracle.gen_frac_golden import write_subpel_yuv  # noqa: E402

RUNS = [  # (W, H, frames, cfg, extra args, cap per (size, phase) class, stride, sub-pel clip?)
    (256, 192, 5, "encoder_randomaccess_main.cfg", ["--SearchRange=16", "-q", "24"], 1, 7, True),
    (416, 240, 3, "encoder_lowdelay_P_main.cfg", ["--SearchRange=32", "-q", "27"], 1, 13, False),
]
HDR = ["magic", "w", "h", "mvx", "mvy", "sad", "pad0", "pad1"]


def parse(path):
    raw = np.fromfile(path, np.int16)
    recs, pos = [], 0
    while pos < raw.size:
        hdr = raw[pos:pos + 16].view(np.int32).copy()
        assert hdr[0] == 0x4d434f53, hex(int(hdr[0]))
        w, h = int(hdr[1]), int(hdr[2])
        pos += 16
        cur = raw[pos:pos + w * h].reshape(h, w).copy()
        pos += w * h
        patch = raw[pos:pos + (w + 8) * (h + 8)].reshape(h + 8, w + 8).copy()
        pos += (w + 8) * (h + 8)
        recs.append((hdr, cur, patch))
    return recs


def parse_ipe(path):
    """Records of xGetInterPredictionError: 12-int header {magic, w, h, lists, had, mv0x, mv0y, mv1x, mv1y, dist, 0, 0}, block, 1 or 2 patches."""
    raw = np.fromfile(path, np.int16)
    recs, pos = [], 0
    while pos < raw.size:
        hdr = raw[pos:pos + 24].view(np.int32).copy()
        assert hdr[0] == 0x49504552, hex(int(hdr[0]))
        w, h, nl = int(hdr[1]), int(hdr[2]), int(hdr[3])
        pos += 24
        cur = raw[pos:pos + w * h].reshape(h, w).copy()
        pos += w * h
        patches = []
        for _ in range(nl):
            patches.append(raw[pos:pos + (w + 8) * (h + 8)].reshape(h + 8, w + 8).copy())
            pos += (w + 8) * (h + 8)
        recs.append((hdr, cur, patches))
    return recs


def main():
    from oracle.pyoracle import Oracle
    binary = os.path.join(REFDIR, "TAppEncoder_cpume")
    recs = []
    ipe = []
    with tempfile.TemporaryDirectory() as d:
        for W, H, F, cfg, extra, cap, stride, subpel in RUNS:
            yuv, log, log2 = os.path.join(d, "c.yuv"), os.path.join(d, "mc.bin"), os.path.join(d, "ipe.bin")
            (write_subpel_yuv if subpel else write_yuv)(yuv, W, H, F)
            for q in (log, log2):
                if os.path.exists(q):
                    os.remove(q)
            env = dict(os.environ, HMME_LOG_MC=log, HMME_LOG_IPE=log2, HMME_LOG_FRAC_CAP=str(cap), HMME_LOG_FRAC_STRIDE=str(stride))
            r = subprocess.run([binary, "-c", os.path.join(REFDIR, "cfg", cfg), "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(F),
                                "-q", "32", "-b", os.path.join(d, "o.hevc"), "-o", os.path.join(d, "rec.yuv"), "--OpenCL=0"] + extra,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
            assert r.returncode == 0, r.stdout[-2000:]
            got = parse(log)
            got2 = parse_ipe(log2)
            print(cfg, W, H, "->", len(got), "template-cost records,", len(got2), "inter-prediction-error records")
            recs += got
            ipe += got2
    O, bad = Oracle(), 0
    for hdr, cur, patch in recs:
        w, h = int(hdr[1]), int(hdr[2])
        # the patch starts 4 samples up-left of the block displaced by the MV's integer part: keep only the fractional part
        pu = np.array([[0, 0, w, h, int(hdr[3]) & 3, int(hdr[4]) & 3]], np.int32)
        sad = int(O.mc_cost(np.ascontiguousarray(cur), (0, 0), np.ascontiguousarray(patch), (4, 4), pu, False)[0])
        if sad != int(np.uint32(hdr[5])):
            bad += 1
            if bad < 10:
                print("MISMATCH", dict(zip(HDR, hdr.tolist())), "oracle", sad)
    print("oracle vs reference records: %d mismatches of %d" % (bad, len(recs)))
    assert bad == 0
    bad = 0
    for hdr, cur, patches in ipe:                     # xGetInterPredictionError: uni- and bi-directional, Hadamard or SAD
        w, h, nl, had = int(hdr[1]), int(hdr[2]), int(hdr[3]), bool(hdr[4])
        if nl == 1:
            pu = np.array([[0, 0, w, h, int(hdr[5]) & 3, int(hdr[6]) & 3]], np.int32)
            got = int(O.mc_cost(np.ascontiguousarray(cur), (0, 0), np.ascontiguousarray(patches[0]), (4, 4), pu, had)[0])
        else:
            pu = np.array([[0, 0, w, h, int(hdr[5]) & 3, int(hdr[6]) & 3, int(hdr[7]) & 3, int(hdr[8]) & 3]], np.int32)
            got = int(O.mc_cost_bi(np.ascontiguousarray(cur), (0, 0), np.ascontiguousarray(patches[0]), np.ascontiguousarray(patches[1]), (4, 4), pu, had)[0])
        if got != int(np.uint32(hdr[9])):
            bad += 1
            if bad < 10:
                print("MISMATCH ipe", hdr.tolist(), "oracle", got)
    print("oracle vs reference inter-prediction-error records: %d mismatches of %d (%d bi-directional)" % (bad, len(ipe), sum(int(r[0][3]) == 2 for r in ipe)))
    assert bad == 0
    out = os.path.join(ROOT, "tests", "golden", "mc_records.npz")
    np.savez_compressed(out, columns=np.array(HDR), hdr=np.stack([r[0] for r in recs]).astype(np.int32),
                        cur=np.concatenate([r[1].ravel() for r in recs]).astype(np.int16),
                        patch=np.concatenate([r[2].ravel() for r in recs]).astype(np.int16),
                        ipe_hdr=np.stack([r[0] for r in ipe]).astype(np.int32),
                        ipe_cur=np.concatenate([r[1].ravel() for r in ipe]).astype(np.int16),
                        ipe_patch=np.concatenate([q.ravel() for r in ipe for q in r[2]]).astype(np.int16))
    hd = np.stack([r[0] for r in recs])
    print("wrote", out, os.path.getsize(out), "bytes;", len(recs), "records;", len(set(map(tuple, hd[:, 1:3].tolist()))), "sizes;",
          len(set(zip((hd[:, 3] & 3).tolist(), (hd[:, 4] & 3).tolist()))), "of 16 quarter-pel phases")


if __name__ == "__main__":
    main()
