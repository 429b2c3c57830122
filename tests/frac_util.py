"""Helpers for the fractional-pel refinement tests: the golden records logged from the
reference encoder (tests/golden/frac_records.npz, made by oracle/gen_frac_golden.py) and a packer that lays them out as
one pair of planes + a PU list, the form the plane-based APIs (oracle and CUDA) take."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["magic", "w", "h", "bi", "had", "mvx", "mvy", "predx", "predy", "lambda", "halfx", "halfy", "qterx", "qtery", "cost", "pad"]


def load_records():
    d = np.load(os.path.join(ROOT, "tests", "golden", "frac_records.npz"))
    hdr, cur, patch, cand = d["hdr"], d["cur"], d["patch"], d["cand"]
    out, pc, pp = [], 0, 0
    for i in range(hdr.shape[0]):
        w, h = int(hdr[i, 1]), int(hdr[i, 2])
        rec = dict(zip(COLS, (int(v) for v in hdr[i])))
        rec["lambda"] = int(np.uint32(hdr[i, 9]))
        rec["cost"] = int(np.uint32(hdr[i, 14]))
        rec["cur"] = cur[pc:pc + w * h].reshape(h, w)
        rec["patch"] = patch[pp:pp + (w + 8) * (h + 8)].reshape(h + 8, w + 8)
        rec["cand"] = cand[i]
        pc += w * h
        pp += (w + 8) * (h + 8)
        out.append(rec)
    return out


def pack_atlas(recs, cell=80, per_row=8, margin=16):
    """Records -> (cur plane, ref plane, origin, pus): record k sits in cell (k % per_row, k // per_row); its reference patch is
    pasted so that the PU at (x, y) with its logged integer MV reads exactly the logged samples.  Cells are 80 apart and MVs
    are applied by displacing the patch, so every record keeps its true MV and predictor."""
    n = len(recs)
    rows = (n + per_row - 1) // per_row
    W, H = per_row * cell, rows * cell
    cur = np.zeros((H + 2 * margin, W + 2 * margin), np.int16)
    pus = np.zeros((n, 8), np.int32)
    # the reference plane cannot hold arbitrary MVs for neighbouring cells at once, so the MV is folded into the predictor:
    # PU k uses integer MV (0,0) and predictor pred - 4*mv, which leaves every bit cost unchanged (the cost only sees differences)
    ref = np.zeros_like(cur)
    for k, r in enumerate(recs):
        x, y = (k % per_row) * cell + 8, (k // per_row) * cell + 8
        w, h = r["w"], r["h"]
        cur[margin + y:margin + y + h, margin + x:margin + x + w] = r["cur"]
        ref[margin + y - 4:margin + y + h + 4, margin + x - 4:margin + x + w + 4] = r["patch"]
        pus[k] = [x, y, w, h, 0, 0, r["predx"] - 4 * r["mvx"], r["predy"] - 4 * r["mvy"]]
    return cur, ref, (margin, margin), pus


def load_mc_records():
    """Records logged from the reference's xGetTemplateCost (tests/golden/mc_records.npz, oracle/gen_mc_golden.py)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "mc_records.npz"))
    hdr, cur, patch = d["hdr"], d["cur"], d["patch"]
    out, pc, pp = [], 0, 0
    for i in range(hdr.shape[0]):
        w, h = int(hdr[i, 1]), int(hdr[i, 2])
        out.append(dict(w=w, h=h, mvx=int(hdr[i, 3]), mvy=int(hdr[i, 4]), sad=int(np.uint32(hdr[i, 5])),
                        cur=cur[pc:pc + w * h].reshape(h, w), patch=patch[pp:pp + (w + 8) * (h + 8)].reshape(h + 8, w + 8)))
        pc += w * h
        pp += (w + 8) * (h + 8)
    return out


def pack_mc_atlas(recs, cell=80, per_row=8, margin=16):
    """Like pack_atlas: record k in its own cell; the patch is pasted at the PU position, so the PU carries only the fractional
    part of the logged MV (its integer part is already folded into where the patch was cut)."""
    n = len(recs)
    rows = (n + per_row - 1) // per_row
    W, H = per_row * cell, rows * cell
    cur = np.zeros((H + 2 * margin, W + 2 * margin), np.int16)
    ref = np.zeros_like(cur)
    pus = np.zeros((n, 6), np.int32)
    for k, r in enumerate(recs):
        x, y = (k % per_row) * cell + 8, (k // per_row) * cell + 8
        w, h = r["w"], r["h"]
        cur[margin + y:margin + y + h, margin + x:margin + x + w] = r["cur"]
        ref[margin + y - 4:margin + y + h + 4, margin + x - 4:margin + x + w + 4] = r["patch"]
        pus[k] = [x, y, w, h, r["mvx"] & 3, r["mvy"] & 3]
    return cur, ref, (margin, margin), pus


def load_ipe_records():
    """Records logged from the reference's xGetInterPredictionError (merge candidates, ME result; uni- and bi-directional)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "mc_records.npz"))
    hdr, cur, patch = d["ipe_hdr"], d["ipe_cur"], d["ipe_patch"]
    out, pc, pp = [], 0, 0
    for i in range(hdr.shape[0]):
        w, h, nl = int(hdr[i, 1]), int(hdr[i, 2]), int(hdr[i, 3])
        r = dict(w=w, h=h, lists=nl, had=int(hdr[i, 4]), mv0=(int(hdr[i, 5]), int(hdr[i, 6])), mv1=(int(hdr[i, 7]), int(hdr[i, 8])),
                 dist=int(np.uint32(hdr[i, 9])), cur=cur[pc:pc + w * h].reshape(h, w), patches=[])
        pc += w * h
        for _ in range(nl):
            r["patches"].append(patch[pp:pp + (w + 8) * (h + 8)].reshape(h + 8, w + 8))
            pp += (w + 8) * (h + 8)
        out.append(r)
    return out


def pack_ipe_atlas(recs, cell=80, per_row=8, margin=16):
    """Records (all with the same number of lists) -> cur plane, one reference plane per list, origin, PU rows (6 or 8 ints)."""
    n, nl = len(recs), recs[0]["lists"]
    rows = (n + per_row - 1) // per_row
    W, H = per_row * cell, rows * cell
    cur = np.zeros((H + 2 * margin, W + 2 * margin), np.int16)
    refs = [np.zeros_like(cur) for _ in range(nl)]
    pus = np.zeros((n, 4 + 2 * nl), np.int32)
    for k, r in enumerate(recs):
        x, y = (k % per_row) * cell + 8, (k // per_row) * cell + 8
        w, h = r["w"], r["h"]
        cur[margin + y:margin + y + h, margin + x:margin + x + w] = r["cur"]
        for l in range(nl):
            refs[l][margin + y - 4:margin + y + h + 4, margin + x - 4:margin + x + w + 4] = r["patches"][l]
        mv = [r["mv0"][0] & 3, r["mv0"][1] & 3] + ([r["mv1"][0] & 3, r["mv1"][1] & 3] if nl == 2 else [])
        pus[k] = [x, y, w, h] + mv
    return cur, refs, (margin, margin), pus
