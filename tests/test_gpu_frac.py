"""GPU parity tests of the fractional-pel refinement, through the C ABI: hmme_refine_frac /
hmme_refine_frame against the CPU oracle and against the records logged from the reference encoder's own
xPatternSearchFracDIF (tests/golden/frac_records.npz).  Bar: bit-exact MVs, costs and all 18 candidate costs."""
import numpy as np
import pytest

from _pkg import hm
from frac_util import load_records, pack_atlas
from synth import frame_jobs, luma_frames, pad_plane

pytestmark = pytest.mark.gpu

SIZES = [(8, 4), (4, 8), (8, 8), (16, 4), (16, 12), (4, 16), (12, 16), (16, 8), (8, 16), (16, 16), (32, 8), (32, 24), (8, 32), (24, 32),
         (32, 16), (16, 32), (32, 32), (64, 16), (64, 48), (16, 64), (48, 64), (64, 32), (32, 64), (64, 64)]


@pytest.fixture(scope="module")
def me():
    m = hm.MotionEstimator(0, 64)
    yield m
    m.close()


def planes(me, cur, ref, W, H, M):
    pc = me.alloc_plane(1 if cur.dtype == np.uint8 else 2, W, H, M, M)
    pr = me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    return pc, pr


def check(res, cand, want, ctx):
    for k, name in ((("mvx", "mvy"), "mvq"),):
        got = np.stack([res[k[0]], res[k[1]]], 1)
        assert np.array_equal(got, want[name]), (ctx, name, np.argwhere(got != want[name])[:5].tolist())
    assert np.array_equal(res["cost"], want["cost"]), (ctx, "cost", np.argwhere(res["cost"] != want["cost"])[:5].tolist())
    assert np.array_equal(res["dist"], want["dist"]), (ctx, "dist")
    if cand is not None:
        assert np.array_equal(cand, want["cand"]), (ctx, "cand", np.argwhere(cand != want["cand"])[:5].tolist())


def test_reference_records_through_c_abi(me):
    """The reference encoder's own inputs and outputs: winners, returned cost and each of the 18 candidate costs."""
    recs = load_records()
    n = 0
    for had in (0, 1):
        for bi in (0, 1):
            for lam in sorted({r["lambda"] for r in recs if r["had"] == had and r["bi"] == bi}):
                sub = [r for r in recs if r["had"] == had and r["bi"] == bi and r["lambda"] == lam]
                cur, ref, (M, _), pus = pack_atlas(sub)
                H, W = cur.shape[0] - 2 * M, cur.shape[1] - 2 * M
                # uni-prediction records go through the 8-bit current plane, bi-prediction ones (2*org - pred) through int16
                pc, pr = planes(me, cur.astype(np.uint8) if not bi else cur, ref.astype(np.uint8), W, H, M)
                me.set_lambda_q16(lam)
                res, cand = me.refine_frac(pc, pr, pus, bool(had), want_candidates=True)
                for k, r in enumerate(sub):
                    assert (int(res["mvx"][k]), int(res["mvy"][k])) == (2 * r["halfx"] + r["qterx"], 2 * r["halfy"] + r["qtery"]), (k, r["w"], r["h"])
                    assert int(res["cost"][k]) == r["cost"]
                    assert np.array_equal(cand[k], r["cand"]), (k, r["w"], r["h"], bi, had)
                n += len(sub)
                pc.free(); pr.free()
    assert n == len(recs)


@pytest.fixture
def kernel_form(request, monkeypatch):
    """Which refinement / distortion kernel a list runs on is the library's choice (list length); the tests pin it, so that every form sees
    every kind of list: "auto" = the library's choice, "group" = the group kernels (four tile processors per warp), "single" = one PU per
    warp (round 1's kernels), "coop" = a CTA per large PU.  The library reads these variables per call."""
    form = request.param
    env = {"auto": {}, "group": {"HMME_FRAC_COOP": "0", "HMME_FRAC_FORM": "2", "HMME_MC_FORM": "2"},
           "single": {"HMME_FRAC_COOP": "0", "HMME_FRAC_FORM": "1", "HMME_MC_FORM": "1"}, "coop": {"HMME_FRAC_COOP": "1"}}[form]
    for k in ("HMME_FRAC_COOP", "HMME_FRAC_FORM", "HMME_MC_FORM"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return form


@pytest.mark.parametrize("kernel_form", ["auto", "group", "single"], indirect=True)
@pytest.mark.parametrize("had,cur16,lam", [(1, 0, 460000), (1, 1, 1000000), (0, 0, 262144), (0, 1, 0), (1, 0, 0), (1, 0, 4500000)])
def test_random_pus_vs_oracle(me, oracle, had, cur16, lam, kernel_form):
    """Every PU size of the 593-partition layout plus odd multiples of 4 (among them two- and three-tile shapes of both Hadamard sizes),
    random integer MVs and predictors, textured content with sub-pel structure; 8-bit and 16-bit (bi-prediction) current planes;
    Hadamard and SAD; on every kernel form."""
    rng = np.random.default_rng(1000 + had * 7 + cur16 * 3 + lam % 97)
    W, H, M = 384, 256, 32
    f = luma_frames(W + 2 * M, H + 2 * M, 2, seed=int(rng.integers(1 << 30)))
    ref = np.ascontiguousarray(f[0].astype(np.int16))
    cur = np.ascontiguousarray(f[1].astype(np.int16))
    if cur16:
        cur = np.ascontiguousarray((2 * cur - rng.integers(0, 256, cur.shape)).astype(np.int16))     # 2*org - pred in [-255, 510]
    sizes = SIZES + [(20, 28), (4, 4), (60, 4), (4, 60), (28, 36), (64, 12), (40, 40), (24, 8), (8, 24), (12, 8), (8, 12), (20, 4), (4, 12), (12, 4)]
    pus = []
    for _ in range(14):
        for (w, h) in sizes:
            x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
            mvx = int(rng.integers(max(-M + 4 - x, -20), min(W + M - 12 - ((w + 7) & ~7) - x, 20) + 1))
            mvy = int(rng.integers(max(-M + 4 - y, -20), min(H + M - 12 - ((h + 7) & ~7) - y, 20) + 1))
            pus.append([x, y, w, h, mvx, mvy, int(rng.integers(-300, 300)), int(rng.integers(-300, 300))])
    pus = np.array(pus, np.int32)
    pc, pr = planes(me, cur.astype(np.uint8) if not cur16 else cur, ref.astype(np.uint8), W, H, M)
    me.set_lambda_q16(lam)
    res, cand = me.refine_frac(pc, pr, pus, bool(had), want_candidates=True)
    want = oracle.refine_frac(cur, (M, M), ref, (M, M), pus, lam, bool(had))
    check(res, cand, want, f"had={had} cur16={cur16} lam={lam}")
    assert len({(int(a), int(b)) for a, b in zip(res["mvx"] - 4 * pus[:, 4], res["mvy"] - 4 * pus[:, 5])}) > 20    # many different winners
    pc.free(); pr.free()


def test_refine_frame_after_search(me, oracle):
    """Whole-frame form: integer search, then all 593 partitions of every CTU refined from the winners left on the device,
    with a per-job predictor; compared with the oracle fed the same integer MVs."""
    W, H, R, M = 256, 128, 16, 40
    f = luma_frames(W, H, 2, seed=42)
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    jobs = frame_jobs(W, H, R)
    lam = 460000
    me.set_lambda_q16(lam)
    pc, pr = planes(me, cur.astype(np.uint8), ref.astype(np.uint8), W, H, M)
    X, Y, _, _ = me.search_frame(pc, pr, jobs, R)
    preds = np.stack([np.arange(len(jobs)) * 3 - 9, 7 - np.arange(len(jobs)) * 2], 1).astype(np.int32)
    for use_had, pp in ((True, preds), (False, None)):
        res = me.refine_frame(pc, pr, len(jobs), pp, use_had)
        rects = me.lib.partition_table()
        pus = np.zeros((len(jobs), 593, 8), np.int32)
        pus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
        pus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
        pus[:, :, 2] = rects[None, :, 2]
        pus[:, :, 3] = rects[None, :, 3]
        pus[:, :, 4], pus[:, :, 5] = X, Y
        if pp is not None:
            pus[:, :, 6], pus[:, :, 7] = pp[:, None, 0], pp[:, None, 1]
        want = oracle.refine_frac(cur, (M, M), ref, (M, M), pus.reshape(-1, 8), lam, use_had)
        check(res.reshape(-1), None, want, f"frame had={use_had}")
    assert me.last_frac_ms() > 0
    pc.free(); pr.free()


def test_frac_errors(me):
    W, H, M = 64, 64, 16
    z = np.zeros((H + 2 * M, W + 2 * M), np.int16)
    pc, pr = planes(me, z.astype(np.uint8), z.astype(np.uint8), W, H, M)
    p16 = me.alloc_plane(2, W, H, M, M)
    ok = np.array([[0, 0, 8, 8, 0, 0, 0, 0]], np.int32)
    me.refine_frac(pc, pr, ok)
    with pytest.raises(hm.HmmeError):          # 16-bit reference plane: not defined for this path
        me.refine_frac(pc, p16, ok)
    for bad in ([0, 0, 6, 8, 0, 0, 0, 0], [0, 0, 8, 68, 0, 0, 0, 0], [0, 0, 0, 8, 0, 0, 0, 0]):
        with pytest.raises(hm.HmmeError):
            me.refine_frac(pc, pr, np.array([bad], np.int32))
    for bad in ([0, 0, 8, 8, -13, 0, 0, 0], [56, 56, 8, 8, 0, 13, 0, 0], [64 + 9, 0, 8, 8, 0, 0, 0, 0]):
        with pytest.raises(hm.HmmeError) as e:
            me.refine_frac(pc, pr, np.array([bad], np.int32))
        assert e.value.code == -6
    with pytest.raises(hm.HmmeError):          # refine_frame without a matching search on this context
        me.refine_frame(pc, pr, 3)
    one = np.array([[0, 0, -12, -12]], np.int32)
    me.search_frame(pc, pr, one, 12)            # window reaches 4 samples from the plane edge: no room for the 8-tap apron
    with pytest.raises(hm.HmmeError) as e:
        me.refine_frame(pc, pr, 1)
    assert e.value.code == -6
    me.search_frame(pc, pr, np.array([[0, 0, -8, -8]], np.int32), 8)
    me.refine_frame(pc, pr, 1)
    pc.free(); pr.free(); p16.free()


def test_refine_pu_host_pointers(me, oracle):
    """hmme_refine_pu: the synchronous host-pointer form the encoder's xPatternSearchFracDIF body maps to."""
    rng = np.random.default_rng(77)
    W, H, M = 128, 96, 24
    f = luma_frames(W + 2 * M, H + 2 * M, 2, seed=9)
    ref = np.ascontiguousarray(f[0].astype(np.int16))
    cur = np.ascontiguousarray((2 * f[1].astype(np.int16) - rng.integers(0, 256, f[1].shape)).astype(np.int16))
    lam = 617503
    me.set_lambda_q16(lam)
    for (w, h) in SIZES:
        x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
        mv = (int(rng.integers(-12, 13)), int(rng.integers(-12, 13)))
        pred = (int(rng.integers(-100, 100)), int(rng.integers(-100, 100)))
        for had in (True, False):
            got = me.refine_pu(cur[M + y:M + y + h, M + x:M + x + w], ref, x, y, M, M, mv, pred, had)
            want = oracle.refine_frac(cur, (M, M), ref, (M, M), np.array([[x, y, w, h, mv[0], mv[1], pred[0], pred[1]]], np.int32), lam, had)
            assert got == (int(want["mvq"][0, 0]), int(want["mvq"][0, 1]), int(want["cost"][0]), int(want["dist"][0])), (w, h, had)
    with pytest.raises(hm.HmmeError) as e:          # 10-bit-like content: not defined for this path
        me.refine_pu(cur[M:M + 8, M:M + 8], ref + 300, 0, 0, M, M, (0, 0), (0, 0))
    assert e.value.code == -5


def test_full_size_1080p_refine(me, oracle):
    """BASELINE configuration: 1080p +-64 search, then all 284 640 PUs refined; every result obeys the size-independent
    properties (offset within +-3 quarter samples, cost >= distortion) and EVERY PU of every CTU equals the oracle."""
    W, H, R, M = 1920, 1080, 64, 80
    f = luma_frames(W, H, 2)
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    jobs = frame_jobs(W, H, R)
    lam = 460000
    me.set_lambda_q16(lam)
    pc, pr = planes(me, cur.astype(np.uint8), ref.astype(np.uint8), W, H, M)
    X, Y, _, cost_int = me.search_frame(pc, pr, jobs, R)
    res = me.refine_frame(pc, pr, len(jobs), None, True)
    assert res.shape == (len(jobs), 593)
    assert np.abs(res["mvx"] - 4 * X).max() <= 3 and np.abs(res["mvy"] - 4 * Y).max() <= 3
    assert (res["cost"] >= res["dist"]).all()
    rects = me.lib.partition_table()
    pus = np.zeros((len(jobs), 593, 8), np.int32)
    pus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
    pus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
    pus[:, :, 2], pus[:, :, 3] = rects[None, :, 2], rects[None, :, 3]
    pus[:, :, 4], pus[:, :, 5] = X, Y
    want = oracle.refine_frac(cur, (M, M), ref, (M, M), pus.reshape(-1, 8), lam, True)
    check(res.reshape(-1), None, want, "1080p, all 284 640 PUs")
    pc.free(); pr.free()


@pytest.mark.parametrize("world", [2, 4])
def test_virtual_bands_refine_equal_whole_frame(me, world):
    """Band sharding (the N-GPU layout of bench.py) leaves the refinement unchanged: each band searches and refines only its
    own jobs; concatenated in rank order the results equal the single-GPU frame."""
    W, H, R, M = 448, 320, 12, 32
    f = luma_frames(W, H, 2, seed=10 + world)
    cur, ref = pad_plane(f[1], M, M, np.uint8), pad_plane(f[0], M, M, np.uint8)
    me.set_lambda_q16(460000)
    pc, pr = planes(me, cur, ref, W, H, M)
    jobs = frame_jobs(W, H, R)
    me.search_frame(pc, pr, jobs, R)
    whole = me.refine_frame(pc, pr, len(jobs), None, True)
    parts = []
    for rank in range(world):
        bj, _ = hm.band_jobs(W, H, R, world, rank)
        if len(bj):
            me.search_frame(pc, pr, bj, R)
            parts.append(me.refine_frame(pc, pr, len(bj), None, True))
    got = np.concatenate(parts, 0)
    for k in ("mvx", "mvy", "cost", "dist"):
        assert np.array_equal(got[k], whole[k]), (world, k)
    pc.free(); pr.free()


def test_mc_cost_reference_records(me):
    """hmme_mc_cost against the SADs the reference encoder's xGetTemplateCost computed (row f3)."""
    from frac_util import load_mc_records, pack_mc_atlas
    recs = load_mc_records()
    cur, ref, (M, _), pus = pack_mc_atlas(recs)
    H, W = cur.shape[0] - 2 * M, cur.shape[1] - 2 * M
    pc, pr = planes(me, cur.astype(np.uint8), ref.astype(np.uint8), W, H, M)
    got = me.mc_cost(pc, pr, pus, False)
    assert np.array_equal(got, np.array([r["sad"] for r in recs], np.uint32)), np.argwhere(got != np.array([r["sad"] for r in recs]))[:5]
    pc.free(); pr.free()


@pytest.mark.parametrize("kernel_form", ["auto", "single"], indirect=True)
@pytest.mark.parametrize("had,cur16", [(0, 0), (1, 0), (1, 1), (0, 1)])
def test_mc_cost_random_vs_oracle(me, oracle, had, cur16, kernel_form):
    """Every PU size, random quarter-pel MVs (all phases, both signs), SAD and Hadamard, 8-bit and 16-bit current planes; the group
    kernel (lists of 64 PUs and more: "auto") and the one-PU-per-warp kernel."""
    rng = np.random.default_rng(300 + 2 * had + cur16)
    W, H, M = 384, 256, 32
    f = luma_frames(W + 2 * M, H + 2 * M, 2, seed=int(rng.integers(1 << 30)))
    ref = np.ascontiguousarray(f[0].astype(np.int16))
    cur = np.ascontiguousarray(f[1].astype(np.int16))
    if cur16:
        cur = np.ascontiguousarray((2 * cur - rng.integers(0, 256, cur.shape)).astype(np.int16))
    pus = []
    for _ in range(12):
        for (w, h) in SIZES + [(20, 28), (4, 4), (60, 4), (64, 12), (40, 40), (24, 8), (8, 24), (12, 8), (8, 12), (20, 4), (4, 12)]:
            x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
            ix = int(rng.integers(max(-M + 4 - x, -20), min(W + M - 12 - ((w + 7) & ~7) - x, 20)))
            iy = int(rng.integers(max(-M + 4 - y, -20), min(H + M - 12 - ((h + 7) & ~7) - y, 20)))
            pus.append([x, y, w, h, 4 * ix + int(rng.integers(0, 4)), 4 * iy + int(rng.integers(0, 4))])
    pus = np.array(pus, np.int32)
    pc, pr = planes(me, cur.astype(np.uint8) if not cur16 else cur, ref.astype(np.uint8), W, H, M)
    got = me.mc_cost(pc, pr, pus, bool(had))
    want = oracle.mc_cost(cur, (M, M), ref, (M, M), pus, bool(had))
    assert np.array_equal(got, want), (had, cur16, np.argwhere(got != want)[:5].tolist())
    with pytest.raises(hm.HmmeError) as e:
        me.mc_cost(pc, pr, np.array([[0, 0, 8, 8, -4 * (M + 1), 0]], np.int32))
    assert e.value.code == -6
    pc.free(); pr.free()


def test_mc_cost_pu_host_pointers(me, oracle):
    """hmme_mc_cost_pu: the synchronous host-pointer form xGetTemplateCost's body maps to."""
    rng = np.random.default_rng(78)
    W, H, M = 128, 96, 24
    f = luma_frames(W + 2 * M, H + 2 * M, 2, seed=19)
    ref = np.ascontiguousarray(f[0].astype(np.int16))
    cur = np.ascontiguousarray(f[1].astype(np.int16))
    for (w, h) in SIZES:
        x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
        mv = (int(rng.integers(-50, 51)), int(rng.integers(-50, 51)))
        for had in (False, True):
            got = me.mc_cost_pu(cur[M + y:M + y + h, M + x:M + x + w], ref, x, y, M, M, mv, had)
            want = int(oracle.mc_cost(cur, (M, M), ref, (M, M), np.array([[x, y, w, h, mv[0], mv[1]]], np.int32), had)[0])
            assert got == want, (w, h, mv, had)


@pytest.mark.parametrize("kernel_form", ["auto", "single"], indirect=True)
@pytest.mark.parametrize("had", [0, 1])
def test_mc_cost_bi_vs_oracle(me, oracle, had, kernel_form):
    """Bi-directional PUs (two reference planes, xPredInterBi + addAvg rounding): plane form and host-pointer form."""
    rng = np.random.default_rng(500 + had)
    W, H, M = 320, 192, 32
    f = luma_frames(W + 2 * M, H + 2 * M, 3, seed=int(rng.integers(1 << 30)))
    ref0, ref1, cur = (np.ascontiguousarray(f[k].astype(np.int16)) for k in (0, 2, 1))
    pus = []
    for _ in range(8):
        for (w, h) in SIZES + [(20, 28), (4, 4), (40, 40), (24, 8), (12, 8), (8, 12)]:
            x, y = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
            mv = []
            for _l in range(2):
                ix = int(rng.integers(max(-M + 4 - x, -20), min(W + M - 12 - ((w + 7) & ~7) - x, 20)))
                iy = int(rng.integers(max(-M + 4 - y, -20), min(H + M - 12 - ((h + 7) & ~7) - y, 20)))
                mv += [4 * ix + int(rng.integers(0, 4)), 4 * iy + int(rng.integers(0, 4))]
            pus.append([x, y, w, h] + mv)
    pus = np.array(pus, np.int32)
    pc = me.alloc_plane(1, W, H, M, M); p0 = me.alloc_plane(1, W, H, M, M); p1 = me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur.astype(np.uint8)); me.upload(p0, ref0.astype(np.uint8)); me.upload(p1, ref1.astype(np.uint8))
    got = me.mc_cost_bi(pc, p0, p1, pus, bool(had))
    want = oracle.mc_cost_bi(cur, (M, M), ref0, ref1, (M, M), pus, bool(had))
    assert np.array_equal(got, want), (had, np.argwhere(got != want)[:5].tolist())
    for k in range(0, len(pus), 7):                   # host-pointer form on a subset
        x, y, w, h, a, b, c_, d = (int(v) for v in pus[k])
        g = me.mc_cost_bi_pu(cur[M + y:M + y + h, M + x:M + x + w], ref0, ref1, x, y, M, M, (a, b), (c_, d), bool(had))
        assert g == int(want[k]), (k, w, h)
    pc.free(); p0.free(); p1.free()


def test_inter_prediction_error_reference_records(me):
    """hmme_mc_cost / hmme_mc_cost_bi against what the reference encoder's xGetInterPredictionError computed."""
    from frac_util import load_ipe_records, pack_ipe_atlas
    recs = load_ipe_records()
    n = 0
    for nl in (1, 2):
        for had in (0, 1):
            sub = [r for r in recs if r["lists"] == nl and r["had"] == had]
            if not sub:
                continue
            cur, refs, (M, _), pus = pack_ipe_atlas(sub)
            H, W = cur.shape[0] - 2 * M, cur.shape[1] - 2 * M
            pc = me.alloc_plane(1, W, H, M, M); me.upload(pc, cur.astype(np.uint8))
            prs = []
            for r_ in refs:
                p_ = me.alloc_plane(1, W, H, M, M); me.upload(p_, r_.astype(np.uint8)); prs.append(p_)
            got = me.mc_cost(pc, prs[0], pus, bool(had)) if nl == 1 else me.mc_cost_bi(pc, prs[0], prs[1], pus, bool(had))
            assert np.array_equal(got, np.array([r["dist"] for r in sub], np.uint32)), (nl, had)
            n += len(sub)
            pc.free()
            for p_ in prs:
                p_.free()
    assert n == len(recs)


@pytest.mark.parametrize("kernel_form", ["auto", "group"], indirect=True)
def test_fuzz_refine_and_mc(me, oracle, kernel_form):
    """Random plane sizes, margins, PU lists (any multiples of 4 up to 64, PUs touching the plane borders so that aprons live in the
    margins), lambdas, modes and batch sizes -- "auto": small batches take the cooperative kernel, large ones the group kernel; "group":
    every list, down to a single PU, on the group kernels (ragged groups, segments with one PU).  Soak with HMME_FUZZ_ITERS / HMME_FUZZ_SEED."""
    import os
    g = np.random.default_rng(int(os.environ.get("HMME_FUZZ_SEED", "2027")))
    for it in range(int(os.environ.get("HMME_FUZZ_ITERS", "40"))):
        W, H = 16 * int(g.integers(5, 20)), 16 * int(g.integers(5, 14))
        M = 16 + 4 * int(g.integers(0, 6))
        cur16 = bool(g.integers(0, 2))
        had = bool(g.integers(0, 2))
        lam = int(g.choice([0, 65536, 460000, 1000000, 4500000, 0xFFFFFFFF]))
        ref = np.ascontiguousarray(g.integers(0, 256, (H + 2 * M, W + 2 * M)).astype(np.int16))
        k = int(g.integers(1, 4))
        ref = np.ascontiguousarray(np.clip((ref + np.roll(ref, k, 0) + np.roll(ref, k, 1) + np.roll(ref, -k, 1)) // 4 * 2 - 100, 0, 255).astype(np.int16))
        cur = np.ascontiguousarray(np.roll(ref, (int(g.integers(-2, 3)), int(g.integers(-2, 3))), (0, 1)) + g.integers(-6, 7, ref.shape).astype(np.int16))
        cur = np.ascontiguousarray((np.clip(cur, 0, 255) if not cur16 else np.clip(2 * cur - 128, -255, 510)).astype(np.int16))
        n = int(g.choice([1, 3, 17, 200, 1500]))
        pus, mcs = [], []
        for _ in range(n):
            w, h = 4 * int(g.integers(1, 17)), 4 * int(g.integers(1, 17))
            x, y = int(g.integers(0, W - w + 1)), int(g.integers(0, H - h + 1))
            w8, h8 = (w + 7) & ~7, (h + 7) & ~7
            mvx = int(g.integers(-M + 4 - x, W + M - 4 - w8 - x + 1))     # the whole legal range, borders included
            mvy = int(g.integers(-M + 4 - y, H + M - 4 - h8 - y + 1))
            mvx, mvy = max(-60, min(60, mvx)), max(-60, min(60, mvy))
            pus.append([x, y, w, h, mvx, mvy, int(g.integers(-400, 400)), int(g.integers(-400, 400))])
            fx, fy = (int(g.integers(0, 4)), int(g.integers(0, 4)))
            ex, ey = min(mvx, W + M - 4 - w8 - x - 1), min(mvy, H + M - 4 - h8 - y - 1)   # a fractional MV reads one more sample
            mcs.append([x, y, w, h, 4 * ex + fx, 4 * ey + fy])
        pus, mcs = np.array(pus, np.int32), np.array(mcs, np.int32)
        pc, pr = planes(me, cur.astype(np.uint8) if not cur16 else cur, ref.astype(np.uint8), W, H, M)
        me.set_lambda_q16(lam)
        res, cand = me.refine_frac(pc, pr, pus, had, want_candidates=True)
        want = oracle.refine_frac(cur, (M, M), ref, (M, M), pus, lam, had)
        check(res, cand, want, f"fuzz {it}: {W}x{H} M={M} n={n} had={had} cur16={cur16} lam={lam}")
        got = me.mc_cost(pc, pr, mcs, had)
        assert np.array_equal(got, oracle.mc_cost(cur, (M, M), ref, (M, M), mcs, had)), f"fuzz {it} mc"
        pc.free(); pr.free()


def test_per_pu_calls_on_a_small_range_context(oracle):
    """A context created for a tiny search range still stages 64x64 PUs (two patches for the bi-directional call)."""
    m = hm.MotionEstimator(0, 1)
    try:
        rng = np.random.default_rng(3)
        M = 16
        ref0 = np.ascontiguousarray(rng.integers(0, 256, (64 + 2 * M, 64 + 2 * M)).astype(np.int16))
        ref1 = np.ascontiguousarray(rng.integers(0, 256, ref0.shape).astype(np.int16))
        cur = np.ascontiguousarray(rng.integers(0, 256, ref0.shape).astype(np.int16))
        m.set_lambda_q16(460000)
        got = m.mc_cost_bi_pu(cur[M:M + 64, M:M + 64], ref0, ref1, 0, 0, M, M, (5, -7), (-9, 14), True)
        want = int(oracle.mc_cost_bi(cur, (M, M), ref0, ref1, (M, M), np.array([[0, 0, 64, 64, 5, -7, -9, 14]], np.int32), True)[0])
        assert got == want
        g2 = m.refine_pu(cur[M:M + 64, M:M + 64], ref0, 0, 0, M, M, (2, -1), (3, 3), True)
        w2 = oracle.refine_frac(cur, (M, M), ref0, (M, M), np.array([[0, 0, 64, 64, 2, -1, 3, 3]], np.int32), 460000, True)
        assert g2 == (int(w2["mvq"][0, 0]), int(w2["mvq"][0, 1]), int(w2["cost"][0]), int(w2["dist"][0]))
    finally:
        m.close()
