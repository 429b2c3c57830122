"""CPU-side checks of the product's boundary: the C-ABI library loads without a GPU, exports exactly what
include/hmme_b200.h declares, and its host-only entry points (layout table) agree with getIndexBlock.
No compute call is made here."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from _pkg import hm
    return hm.HmmeLib.get()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hmme_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hmme_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib):
    from _pkg import hm
    from importlib import import_module
    api = import_module("hm_opencl_b200.api")
    declared = header_symbols()
    assert declared == sorted(api.EXPORTS)
    nm = subprocess.run(["nm", "-D", "--defined-only", hm.lib_path()], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (hmme_[a-z0-9_]+)", nm)))
    assert exported == declared


def test_library_carries_sm100a_code_only(lib):
    from _pkg import hm
    out = subprocess.run(["cuobjdump", "-lelf", hm.lib_path()], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out)


def test_product_layout_matches_getindexblock(lib):
    from hevc_geom import decode_key, pu_rect
    cases = json.load(open(os.path.join(ROOT, "tests/golden/getindexblock_593.json")))["cases"]
    tab = lib.partition_table()
    for key, idx in cases:
        assert tuple(tab[idx]) == pu_rect(*decode_key(key)), (key, idx)


def test_closed_form_index_block_equals_the_reference_switch(lib):
    """hmme_index_block must return the reference's index for all 593 listed keys and -1 for every other key the
    switch does not list (checked exhaustively over the whole key space the encoder can form)."""
    from hevc_geom import decode_key
    cases = dict((k, v) for k, v in json.load(open(os.path.join(ROOT, "tests/golden/getindexblock_593.json")))["cases"])
    for key, idx in cases.items():
        w, h, z, ps, depth, part = decode_key(key)
        assert lib.index_block(ps, depth, part, z, w, h) == idx, key
    hits = 0
    for depth in range(4):
        S = 64 >> depth
        for ps in range(8):
            for part in range(2):
                for z in range(256):
                    key = S + 100 * (S + 100 * (z + 1000 * (ps + 10 * depth + 100 * part)))
                    got = lib.index_block(ps, depth, part, z, S, S)
                    assert got == cases.get(key, -1), (depth, ps, part, z)
                    hits += got >= 0
    assert hits == 593
    assert lib.index_block(0, 0, 0, 0, 32, 64) == -1 and lib.index_block(0, 4, 0, 0, 4, 4) == -1


def test_search_window_matches_the_reference_encoder(lib):
    """hmme_search_window against every window placement the reference encoder itself computed (xSetSearchRange + clipMv,
    logged by the instrumented build; oracle/gen_window_golden.py): uni- and bi-prediction calls, ranges 4/8/64/96,
    clipping at all four picture borders."""
    g = json.load(open(os.path.join(ROOT, "tests/golden/search_window_lt.json")))
    rows = g["rows"]
    assert len(rows) > 100
    ranges, clipped = set(), 0
    for ph, pv, R, cx, cy, W, H, ltx, lty, rbx, rby in rows:
        assert lib.search_window(ph, pv, R, cx, cy, W, H) == (ltx, lty, rbx, rby), (ph, pv, R, cx, cy)
        ranges.add(R)
        clipped += (rbx - ltx != 2 * R) or (rby - lty != 2 * R)
    assert {4, 8, 64, 96} <= ranges and clipped > 10


def test_python_mirror_constants_match_the_header(lib):
    """The ctypes mirror repeats a few constants of include/hmme_b200.h: frame slots of a group, partitions per CTU, distribution modes."""
    from _pkg import hm
    src = open(os.path.join(ROOT, "include", "hmme_b200.h")).read()
    assert int(re.search(r"#define\s+HMME_GROUP_SLOTS\s+(\d+)", src).group(1)) == hm.Group.SLOTS
    assert int(re.search(r"#define\s+HMME_NUM_CTU_PARTS\s+(\d+)", src).group(1)) == 593 == len(lib.partition_table())
    m = re.search(r"HMME_REF_BAND_HALO\s*=\s*(\d+)\s*,\s*HMME_REF_BROADCAST\s*=\s*(\d+)", src)
    assert (int(m.group(1)), int(m.group(2))) == (hm.Group.BAND_HALO, hm.Group.BROADCAST)


def test_product_and_oracle_layout_agree(lib, oracle):
    assert np.array_equal(lib.partition_table(), oracle.partition_table())


def test_no_gpu_is_a_loud_error(lib):
    import torch
    from _pkg import hm
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hm.HmmeError):
        hm.MotionEstimator(0, 64)
    t = hm.TEncOpenCL()
    assert t.findDevice(0) is False


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the product may import, link or name it."""
    pkg = os.path.join(ROOT, "hm-opencl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "pyoracle" not in txt and "hmme_oracle" not in txt and "oracle/" not in txt.replace("never imported here", "").replace("The oracle under oracle/", ""), f
