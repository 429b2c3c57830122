#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python oracle/gen_golden.py

Writes
  tests/golden/getindexblock_593.json  -- every `case K: index = V` of the active (#else, 593-entry)
        switch of TComDataCU::getIndexBlock (/root/reference/source/Lib/TLibCommon/TComDataCU.cpp:4676-6461)
  tests/golden/refemu_vectors.npz      -- inputs + outputs of the reference's own
        TEncOpenCL::calcMotionVectors (TEncOpenCL.cpp:240-362) driving its own cl/sad.cl kernels in
        lock-step on the CPU (oracle/refemu/), for the edge cases SURVEY.md section 4 lists.

TEST INFRASTRUCTURE ONLY.  The GPU box has no /root/reference: tests read the fixtures, never this script.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefEmu, build  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def parse_getindexblock():
    src = open(os.path.join(REF, "source/Lib/TLibCommon/TComDataCU.cpp")).read().split("\n")
    # the function starts at :3379; the 593-entry variant is the #else branch of `#if AMP_ENC_SPEEDUP`
    start = next(i for i, l in enumerate(src) if "TComDataCU::getIndexBlock" in l)
    els = next(i for i in range(start, len(src)) if src[i].strip().startswith("#else"))
    end = next(i for i in range(els, len(src)) if src[i].strip().startswith("#endif"))
    body = "\n".join(src[els:end])
    cases = re.findall(r"case\s+(\d+)\s*:\s*index\s*=\s*(\d+)\s*;", body)
    return [[int(k), int(v)] for k, v in cases]


def make_cases():
    """(name, cur int16 64x64, plane int16, margin, R, ltx, lty, lambda_q16)"""
    out = []
    g = np.random.default_rng(20261018)

    def plane_u8(R, extra=0):
        m = R + 2 + extra
        return g.integers(0, 256, size=(64 + 2 * m + 64, 64 + 2 * m + 64)).astype(np.int16), m

    # 1. i.i.d. random, centred window, several R and lambda
    for R, lam in [(1, 0), (4, 262144), (8, 460000), (16, 1000000)]:
        p, m = plane_u8(R)
        cur = g.integers(0, 256, size=(64, 64)).astype(np.int16)
        out.append((f"random_R{R}", cur, p, m, R, -R, -R, lam))
    # 2. translated content (true motion inside the window) + noise, off-centre LT (clipped window case)
    p, m = plane_u8(8, extra=6)
    cur = p[m + 3:m + 67, m - 5:m + 59].copy()
    cur = np.clip(cur + g.integers(-6, 7, size=cur.shape), 0, 255).astype(np.int16)
    out.append(("shifted_offcentre_R8", cur, p, m, 8, -13, -3, 460000))
    # 3. constant planes: every candidate ties on SAD -> tie-break + bit-cost ordering
    p = np.full((64 + 40, 64 + 40), 77, np.int16)
    out.append(("constant_ties_lam0_R6", np.full((64, 64), 77, np.int16), p, 20, 6, -6, -6, 0))
    out.append(("constant_ties_lam_R6", np.full((64, 64), 90, np.int16), p, 20, 6, -6, -6, 460000))
    # 4. single impulse in the reference
    p = np.zeros((64 + 40, 64 + 40), np.int16)
    p[20 + 30, 20 + 17] = 255
    cur = np.zeros((64, 64), np.int16)
    cur[28, 20] = 255
    out.append(("impulse_R5", cur, p, 20, 5, -5, -5, 262144))
    # 5. gradient (monotone SAD surface)
    yy, xx = np.mgrid[0:64 + 40, 0:64 + 40]
    p = ((xx * 2 + yy * 3) % 256).astype(np.int16)
    cur = p[20 + 2:20 + 66, 20 + 1:20 + 65].copy()
    out.append(("gradient_R4", cur, p, 20, 4, -4, -4, 4500000))
    # 6. bi-prediction refinement call: cur = 2*org - pred in [-255, 510], R = 4 (TEncSearch.cpp:3702-3712)
    p, m = plane_u8(4)
    org = g.integers(0, 256, size=(64, 64)).astype(np.int32)
    pred = g.integers(0, 256, size=(64, 64)).astype(np.int32)
    out.append(("bipred_16bit_R4", (2 * org - pred).astype(np.int16), p, m, 4, -2, -6, 460000))
    # 7. extreme contrast (max 4x4 SAD = 4080, 64x64 = 1044480) and a huge lambda (32-bit wrap of lambda*bits)
    p = np.full((64 + 24, 64 + 24), 255, np.int16)
    p[::7, ::5] = 0
    out.append(("max_contrast_wrap_lambda_R3", np.zeros((64, 64), np.int16), p, 12, 3, -3, -3, 0xF0000000))
    # 8. positive-only window (LT > 0) so mv bits take the positive branch everywhere
    p, m = plane_u8(4, extra=12)
    cur = g.integers(0, 256, size=(64, 64)).astype(np.int16)
    out.append(("positive_window_R4", cur, p, m, 4, 3, 5, 460000))
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (run in the build container)")
    build(ref=True)
    os.makedirs(GOLD, exist_ok=True)
    cases = parse_getindexblock()
    assert len(cases) == 593, len(cases)
    with open(os.path.join(GOLD, "getindexblock_593.json"), "w") as f:
        json.dump({"source": "TComDataCU.cpp getIndexBlock, #else branch (AMP_ENC_SPEEDUP=0)",
                   "key": "w + 100*(h + 100*(zIdx + 1000*(partSize + 10*depth + 100*partIdx)))",
                   "cases": cases}, f)
    ref = RefEmu(os.path.join(REF, "cl/sad.cl"), 16)
    store = {}
    names = []
    for name, cur, plane, m, R, ltx, lty, lam in make_cases():
        ref.set_lambda_q16(lam)
        X, Y, S, Cst = ref.calc(cur, plane, 0, 0, m, m, R, ltx, lty)
        names.append(name)
        store[name + ".cur"] = cur
        store[name + ".plane"] = plane
        store[name + ".meta"] = np.array([m, R, ltx, lty, lam], np.int64)
        store[name + ".X"], store[name + ".Y"], store[name + ".sad"], store[name + ".cost"] = X, Y, S, Cst
        print(f"{name:32s} R={R} lt=({ltx},{lty}) lam={lam}  mv[592]=({X[592]},{Y[592]}) sad={S[592]} cost={Cst[592]}")
    # lambda quantisation through the reference's own setLambda (TEncOpenCL.h:121)
    lams = [0.0, 1.0, 16.0, 49.3, 57.908390375799, 1234.5678, 4700.0]
    store["lambda.in"] = np.array(lams, np.float64)
    store["lambda.q16"] = np.array([ref.set_lambda(v) for v in lams], np.uint32)
    store["names"] = np.array(names)
    print("refemu stats:", ref.stats())
    ref.close()
    np.savez_compressed(os.path.join(GOLD, "refemu_vectors.npz"), **store)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
