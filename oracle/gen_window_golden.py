#!/usr/bin/env python
"""Golden vectors for the search-window placement FROM THE REFERENCE ITSELF.

The instrumented reference encoder (oracle/_ref/TAppEncoder_cpume, see patch_cpume.py) prints, for every
calcMotionVectors call, the MV it centres the window on, the range, the CU position, the picture size and the LT/RB
it computed with TEncSearch::xSetSearchRange (TEncSearch.cpp:3814-3830) + TComDataCU::clipMv (TComDataCU.cpp:2907-2920).
Unique tuples go to tests/golden/search_window_lt.json; tests compare hmme_search_window() against them.
Needs /root/reference and `make -C oracle encoders`.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_encoder_golden import REFDIR, write_yuv  # noqa: E402

RUNS = [  # (W, H, frames, cfg, range)
    (416, 240, 5, "encoder_randomaccess_main.cfg", 8),
    (416, 240, 3, "encoder_lowdelay_P_main.cfg", 64),
    (192, 128, 4, "encoder_lowdelay_P_main.cfg", 96),
]


def main():
    binary = os.path.join(REFDIR, "TAppEncoder_cpume")
    rows = set()
    with tempfile.TemporaryDirectory() as d:
        for W, H, F, cfg, R in RUNS:
            yuv = os.path.join(d, "c.yuv")
            write_yuv(yuv, W, H, F)
            env = dict(os.environ, HMME_LOG_LT="1")
            r = subprocess.run([binary, "-c", os.path.join(REFDIR, "cfg", cfg), "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(F),
                                "-q", "32", "-b", os.path.join(d, "o.hevc"), "-o", os.path.join(d, "rec.yuv"), "--OpenCL=1", "--KernelOpenCL=/root/reference/cl/sad.cl",
                                "--SearchRange=%d" % R], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
            assert r.returncode == 0, r.stdout[-2000:]
            for line in r.stdout.splitlines():
                if line.startswith("HMME_LT "):
                    rows.add(tuple(int(v) for v in line.split()[1:]))
            print(cfg, W, H, R, "->", len(rows), "unique placements so far")
    out = sorted(rows)
    with open(os.path.join(ROOT, "tests", "golden", "search_window_lt.json"), "w") as f:
        json.dump({"columns": ["predHorQpel", "predVerQpel", "range", "cuX", "cuY", "picW", "picH", "ltx", "lty", "rbx", "rby"],
                   "source": "TEncSearch::xSetSearchRange + TComDataCU::clipMv as executed by the reference encoder (oracle/gen_window_golden.py)",
                   "rows": out}, f)
    print("wrote", len(out), "rows")


if __name__ == "__main__":
    main()
