"""Runs bench.py (short) on the GPU and checks the JSON line against the contract the driver reads."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_on_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "12", "--warmup", "3", "--no-cpu-baseline"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["metric"] == "me_block_sad_evaluations_per_s" and d["unit"] == "block-SAD evaluations/s" and d["n_gpus"] == 1
    assert d["steps"] == 12 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 480 * 16641 * 593 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    assert d["gpu_launches"] == 2 * 12                              # search + finalize per step, all of them this library's kernels
    e = d["e2e"]
    # int16 host planes: the band's CTU rows of the current frame (1920 x 1024) + band and halo of the reference (2048 x 1152), jobs
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * (1920 * 1024 + 2048 * 1152) + 480 * 16 and e["d2h_bytes_per_step"] == 480 * 593 * 16
    assert e["band_halo_u8"]["value"] > 0
    v = d["verified"]
    assert v["ctus"] == 480 and v["mismatches"] == 0 and v["ctu_result_sets_compared"] == (d["e2e"]["slots"] + 1) * 480
    rf = d["roofline"]
    assert rf["bound"] == "int_issue" and 0.3 < rf["frac"] < 1.0 and rf["peak"] > 20 and rf["achieved"] > 5 and rf["traffic"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and abs(rf["frac_issue"] - rf["frac"]) < 1e-9 and abs(rf["frac_one_pipe"] - 2 * rf["frac"]) < 1e-9
    assert d["frac_refine"]["roofline"]["frac_issue"] > 0
    pc = d["per_ctu"]
    assert pc["calls"] == 480 and 0 < pc["latency_ms"] < 5
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert 0.5 < d["ms_per_step"] < 5.0                             # 1080p +-64 on one B200: ~1.3 ms


def test_random_access_workload_verifies_every_table():
    """BASELINE config[2] at full size: two reference lists + the bi-prediction refinement of every CTU of a 1080p B frame, results in a
    device-resident table, all 3 x 480 result sets compared with the oracle inside bench.py."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "1080p64_ra", "--steps", "6", "--warmup", "3", "--no-cpu-baseline"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    ra = d["random_access"]
    assert ra["verified"] == {"ctus": 1440, "mismatches": 0, "what": ra["verified"]["what"]}
    assert 1.0 < ra["ms_per_b_frame"] < 10.0 and 0 < ra["bipred_kernel_ms"] < 1.0 and ra["bipred_roofline"]["frac"] > 0
