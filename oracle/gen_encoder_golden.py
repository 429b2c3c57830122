#!/usr/bin/env python
"""Whole-encoder golden bitstreams FROM THE REFERENCE ITSELF.

Runs oracle/_ref/TAppEncoder_refcl -- every reference source unmodified, its own TEncOpenCL.cpp + cl/sad.cl
GPU-ME path executing on the CPU behind oracle/refemu (lock-step) -- on synthetic clips and records the MD5 of
the bitstream and of the reconstruction in tests/golden/encoder_bitstreams.json.  tests/test_gpu_encoder.py
replays the same command lines with oracle/_ref/TAppEncoder_b200 (the same encoder with TEncOpenCL swapped for
hm-opencl_b200/host/ + libhmme_b200.so) on a B200 and requires identical MD5s.

Needs /root/reference and `make -C oracle encoders`.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synth import luma_frames  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")

# name: (W, H, frames, cfg, extra args)
CASES = {
    "ldp_416x240_4f_sr8": (416, 240, 4, "encoder_lowdelay_P_main.cfg", ["--SearchRange=8"]),
    "ldp_416x240_2f_sr64": (416, 240, 2, "encoder_lowdelay_P_main.cfg", ["--SearchRange=64"]),
    "ra_416x240_9f_sr8": (416, 240, 9, "encoder_randomaccess_main.cfg", ["--SearchRange=8"]),
    "ldp_1920x1080_2f_sr8": (1920, 1080, 2, "encoder_lowdelay_P_main.cfg", ["--SearchRange=8"]),
    "ra_416x240_9f_sr64": (416, 240, 9, "encoder_randomaccess_main.cfg", ["--SearchRange=64"]),
    # BASELINE.json configs[1] at its full size: 480 calcMotionVectors calls of 16 641 candidates each (~35 min of lock-step emulation)
    "ldp_1920x1080_2f_sr64": (1920, 1080, 2, "encoder_lowdelay_P_main.cfg", ["--SearchRange=64"]),
    # BASELINE.json configs[2] at its full picture size: random access (two lists + bi-prediction refinement), 3 pictures (~100 min of emulation)
    "ra_1920x1080_3f_sr64": (1920, 1080, 3, "encoder_randomaccess_main.cfg", ["--SearchRange=64"]),
}


def write_yuv(path, W, H, F):
    with open(path, "wb") as f:
        for y in luma_frames(W, H, F):
            f.write(y.tobytes())
            f.write(np.full((H // 2) * (W // 2) * 2, 128, np.uint8).tobytes())


def encoder_args(W, H, F, cfg, extra, yuv, bit, rec, kernel_path):
    return ["-c", os.path.join(REFDIR, "cfg", cfg), "-i", yuv, "-wdt", str(W), "-hgt", str(H), "-fr", "30", "-f", str(F), "-q", "32",
            "-b", bit, "-o", rec, "--OpenCL=1", "--OpenCLDevice=0", "--KernelOpenCL=" + kernel_path, "--SEIDecodedPictureHash=1"] + extra


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def decode_check(bitstream, recon, workdir):
    """Reference decoder (oracle/_ref/TAppDecoder_ref): every picture's decoded-picture-hash SEI must verify "(OK)" and the
    decoder output must equal the encoder's reconstruction.  Returns the number of pictures checked, or None if the
    decoder was not built."""
    dec = os.path.join(REFDIR, "TAppDecoder_ref")
    if not os.path.exists(dec):
        return None
    out = os.path.join(workdir, "dec.yuv")
    r = subprocess.run([dec, "-b", bitstream, "-o", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "***ERROR***" not in r.stdout, r.stdout[-2000:]
    assert md5(out) == md5(recon), "decoder output differs from the encoder reconstruction"
    return r.stdout.count("(OK)")


def run_case(binary, name, kernel_path, workdir, env=None):
    W, H, F, cfg, extra = CASES[name]
    yuv, bit, rec = (os.path.join(workdir, name + e) for e in (".yuv", ".hevc", "_rec.yuv"))
    write_yuv(yuv, W, H, F)
    t0 = time.time()
    r = subprocess.run([binary] + encoder_args(W, H, F, cfg, extra, yuv, bit, rec, kernel_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       env=dict(os.environ, **env) if env else None)
    if r.returncode != 0:
        raise RuntimeError("encoder failed:\n" + r.stdout[-3000:])
    return {"bitstream_md5": md5(bit), "bitstream_bytes": os.path.getsize(bit), "recon_md5": md5(rec), "yuv_md5": md5(yuv),
            "decoded_ok": decode_check(bit, rec, workdir), "frames": F,
            "seconds": round(time.time() - t0, 1), "poc_lines": [l.strip()[:120] for l in r.stdout.splitlines() if l.startswith("POC")],
            "spec_line": next((l.strip() for l in r.stdout.splitlines() if l.startswith("HMME_SPEC")), None)}


def main():
    binary = os.path.join(REFDIR, "TAppEncoder_refcl")
    if not os.path.exists(binary):
        sys.exit("build it first: make -C oracle encoders")
    gold_path = os.path.join(ROOT, "tests", "golden", "encoder_bitstreams.json")
    out = json.load(open(gold_path))["cases"] if os.path.exists(gold_path) and "--all" not in sys.argv else {}
    with tempfile.TemporaryDirectory() as d:
        for name in CASES:
            if name in out:
                continue                       # lock-step emulation is slow: only new cases are generated unless --all
            out[name] = run_case(binary, name, "/root/reference/cl/sad.cl", d)
            print(name, out[name]["bitstream_md5"], out[name]["bitstream_bytes"], "bytes", out[name]["seconds"], "s")
    with open(gold_path, "w") as f:
        json.dump({"generator": "oracle/gen_encoder_golden.py with oracle/_ref/TAppEncoder_refcl (reference sources + lock-step OpenCL emulation)",
                   "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
