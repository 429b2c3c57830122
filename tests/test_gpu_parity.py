"""GPU parity tests (run by the driver with -m gpu on a B200): the CUDA path, called through the C ABI
(libhmme_b200.so via ctypes), against the CPU oracle and the reference-derived golden vectors.
Bar: bit-exact X, Y, sad (ruiCosts) and cost (minSad) -- integer work."""
import os

import numpy as np
import pytest

from _pkg import hm
from synth import frame_jobs, luma_frames, pad_plane

pytestmark = pytest.mark.gpu
NAMES = ("X", "Y", "sad", "cost")


@pytest.fixture(scope="module")
def me():
    m = hm.MotionEstimator(0, 128)
    yield m
    m.close()


def assert_same(got, want, ctx=""):
    for g, w, n in zip(got, want, NAMES):
        if not np.array_equal(g, w):
            bad = np.argwhere(np.asarray(g) != np.asarray(w))
            raise AssertionError(f"{ctx}: {n} differs at {bad[:8].tolist()} got {np.asarray(g)[tuple(bad[0])]} want {np.asarray(w)[tuple(bad[0])]} ({len(bad)} mismatches)")


def test_golden_vectors_through_c_abi(me, golden):
    """Every reference-derived fixture (8-bit, 16-bit bi-pred cur, ties, wrap-around lambda ...) via hmme_search_ctu."""
    for name in [str(n) for n in golden["names"]]:
        m, R, ltx, lty, lam = (int(v) for v in golden[name + ".meta"])
        me.set_lambda_q16(lam)
        got = me.search_ctu(golden[name + ".cur"], golden[name + ".plane"], 0, 0, m, m, R, ltx, lty)
        assert_same(got, [golden[f"{name}.{k}"] for k in NAMES], name)


@pytest.mark.parametrize("R", [1, 4, 7, 16, 32])
def test_frame_batch_vs_oracle(me, oracle, R):
    W, H = 256, 128
    f = luma_frames(W, H, 2, seed=100 + R)
    M = R + 16
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    jobs = frame_jobs(W, H, R)
    jobs[1::2, 2] += 5      # off-centre windows on every other CTU
    jobs[1::2, 3] -= 3
    lam = 460000
    me.set_lambda_q16(lam)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8)
    assert_same(got, want, f"R={R}")
    pc.free(); pr.free()


@pytest.mark.parametrize("R,lam", [(64, 460000), (64, 0), (128, 1000000)])
def test_headline_ranges_vs_oracle(me, oracle, R, lam):
    """+-64 (BASELINE config 2) and +-128 (config 4) on a few CTUs: multi-tile merge, 9 and 2x43 tiles per job."""
    W, H = 128, 64
    f = luma_frames(W, H, 2, seed=7 + R)
    M = R + 8
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    jobs = frame_jobs(W, H, R)
    me.set_lambda_q16(lam)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8)
    assert_same(got, want, f"R={R}")
    pc.free(); pr.free()


def test_all_ties_tiebreak_across_tiles(me, oracle):
    """Constant planes: every candidate has SAD 0; with lambda = 0 the first candidate in scan order (LT) must win
    in every partition even though 9 tiles race through atomicMin; with lambda > 0 the cheapest MV wins."""
    R, W, H = 64, 64, 64
    M = R + 8
    cur = np.full((H + 2 * M, W + 2 * M), 93, np.int16)
    ref = cur.copy()
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = np.array([[0, 0, -R, -R]], np.int32)
    for lam in (0, 460000):
        me.set_lambda_q16(lam)
        got = me.search_frame(pc, pr, jobs, R)
        want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam)
        assert_same(got, want, f"ties lam={lam}")
        exp = -R if lam == 0 else 0
        assert (got[0] == exp).all() and (got[1] == exp).all()
    pc.free(); pr.free()


def test_extreme_contrast_64(me, oracle):
    """cur = 0, ref = 255 everywhere: 4x4 SAD = 4080, 64x64 SAD = 1044480 (largest value the packed keys must hold),
    together with the largest lambda that keeps mvcost at 16 bits."""
    R, W, H = 16, 64, 64
    M = R + 8
    cur = np.zeros((H + 2 * M, W + 2 * M), np.int16)
    ref = np.full_like(cur, 255)
    ref[M + 70, M + 3] = 254
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = np.array([[0, 0, -R, -R]], np.int32)
    for lam in (0, 0xFFFFFFFF, 0x7FFF0000):
        me.set_lambda_q16(lam)
        got = me.search_frame(pc, pr, jobs, R)
        want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam)
        assert_same(got, want, f"contrast lam={lam:#x}")
    pc.free(); pr.free()


@pytest.mark.parametrize("ce,re_", [(2, 1), (1, 2), (2, 2)])
def test_16bit_planes_generic_kernel(me, oracle, ce, re_):
    """Bi-prediction style input: cur = 2*org - pred (signed 16 bit), R = 4 (bipredSearchRange), and 16-bit reference planes."""
    R, W, H = 4, 128, 64
    g = np.random.default_rng(3)
    M = R + 8
    shape = (H + 2 * M, W + 2 * M)
    cur = (g.integers(-255, 511, size=shape) if ce == 2 else g.integers(0, 256, size=shape)).astype(np.int16)
    ref = (g.integers(-300, 700, size=shape) if re_ == 2 else g.integers(0, 256, size=shape)).astype(np.int16)
    pc, pr = me.alloc_plane(ce, W, H, M, M), me.alloc_plane(re_, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = frame_jobs(W, H, R)
    me.set_lambda_q16(460000)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, 460000)
    assert_same(got, want, f"elem {ce},{re_}")
    pc.free(); pr.free()


def test_row_wrap_linear_addressing(me, oracle):
    """A window pushed past the right margin reads on into the next row of the padded plane:
    the CUDA path must reproduce the reference's linear addressing."""
    R, W, H = 8, 128, 128
    f = luma_frames(W, H, 2, seed=5)
    M = 16
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    assert pr.desc.pitch == ref.shape[1]       # 128 + 32 = 160 is already a multiple of 16: same linear layout as the host plane
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = np.array([[64, 0, 10, -4], [64, 64, 12, -20]], np.int32)   # x: 64+10+2*8+63 = 153 > 128+16
    me.set_lambda_q16(262144)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, 262144)
    assert_same(got, want, "row wrap")
    pc.free(); pr.free()


def test_errors_are_loud(me):
    W, H, M = 64, 64, 8
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    ok = np.zeros((H + 2 * M, W + 2 * M), np.int16)
    me.upload(pc, ok); me.upload(pr, ok)
    with pytest.raises(hm.HmmeError) as e:
        me.search_frame(pc, pr, np.array([[0, 0, -16, -16]], np.int32), 16)     # window leaves the allocation
    assert e.value.code == -6
    bad = ok.copy(); bad[3, 3] = 300
    with pytest.raises(hm.HmmeError) as e:
        me.upload(pr, bad)                                                       # not 8-bit content
    assert e.value.code == -5
    with pytest.raises(hm.HmmeError) as e:
        me.search_ctu(ok[:64, :64], ok, 0, 0, M, M, 4096, 0, 0)                  # beyond createBuffers' range
    assert e.value.code == -4
    with pytest.raises(hm.HmmeError):
        hm.MotionEstimator(0, 64).__class__(999, 64)                             # no such device
    pc.free(); pr.free()


def test_python_mirror_of_reference_class(oracle, golden):
    """Same call order as TEncTop::xInitOpenCL + TEncSearch::xMotionEstimation."""
    t = hm.TEncOpenCL()
    assert t.findDevice(0)
    assert not t.compileKernelSource("cl/sad.cl", "calcSAD")       # 425-entry kernel is not part of the path
    assert t.compileKernelSource("cl/sad.cl", "calcSAD_AMP")
    assert t.createBuffers(64, 64, 64)
    with pytest.raises(hm.HmmeError):
        t.calcMotionVectors(np.zeros((64, 64), np.int16), np.zeros((200, 200), np.int16), (70, 70), 4, (-4, -4))
    t.setEnabled(True)
    t.setLambda(49.3)
    name = "shifted_offcentre_R8"
    m, R, ltx, lty, _ = (int(v) for v in golden[name + ".meta"])
    t.calcMotionVectors(golden[name + ".cur"], golden[name + ".plane"], (m, m), R, (ltx, lty))
    lam = oracle.lambda_q16(49.3)
    want = oracle.search_ctu(golden[name + ".cur"], golden[name + ".plane"], 0, 0, m, m, R, ltx, lty, lam)
    assert np.array_equal(t.getX(), want[0]) and np.array_equal(t.getY(), want[1]) and np.array_equal(t.getRuiCost(), want[2])


def test_full_1080p_pm64_every_ctu(me, oracle):
    """BASELINE config 2 at full size: 480 CTUs x 16641 candidates.  Size-independent properties on every job
    (global pan recovered by the 64x64 partition, cost = sad + mvcost(winner), winners inside the window) and a
    bit-exact comparison of EVERY CTU with the oracle (the threaded oracle does the frame in about a second)."""
    W, H, R = 1920, 1080, 64
    f = luma_frames(W, H, 2)
    M = 80 + 64
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = frame_jobs(W, H, R)
    assert len(jobs) == 480
    lam = 460000
    me.set_lambda_q16(lam)
    X, Y, S, Cst = me.search_frame(pc, pr, jobs, R)
    assert (X >= -R).all() and (X <= R).all() and (Y >= -R).all() and (Y <= R).all()
    bits = np.vectorize(lambda v: oracle.mv_bits(4 * int(v)))
    mvc = ((lam * (bits(X) + bits(Y)).astype(np.uint64)) & 0xFFFFFFFF) >> 16
    assert np.array_equal(Cst.astype(np.uint64), S.astype(np.uint64) + mvc)
    # frame 1 is frame 0 panned by (+3,+2): away from the moving square the 64x64 block matches at MV (3,2) with SAD 0
    hit = (X[:, 592] == 3) & (Y[:, 592] == 2) & (S[:, 592] == 0)
    assert hit.sum() >= 450      # all CTUs except those under the moving square or whose match leaves the picture
    # hierarchy consistency where winners coincide: the 64x64 SAD is the sum of the four 32x32 SADs
    same = hit & np.all(X[:, 584:588] == 3, 1) & np.all(Y[:, 584:588] == 2, 1)
    assert same.sum() >= 400 and (S[same, 584:588].sum(1) == S[same, 592]).all()
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=os.cpu_count() or 8)
    assert_same((X, Y, S, Cst), want, "1080p, all 480 CTUs")
    pc.free(); pr.free()


def test_full_4k_pm128_every_ctu(me, oracle):
    """BASELINE config 4 geometry at full size on one GPU: 3840x2160, +-128, 1980 CTUs x 66049 candidates (2 x 43 tiles per
    job).  Properties on every job + bit-exact comparison of EVERY CTU with the oracle (~10-20 s on the box's cores)."""
    W, H, R = 3840, 2160, 128
    f = luma_frames(W, H, 2, seed=77)
    M = R + 16
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = frame_jobs(W, H, R)
    assert len(jobs) == 1980
    lam = 1000000
    me.set_lambda_q16(lam)
    X, Y, S, Cst = me.search_frame(pc, pr, jobs, R)
    assert (X >= -R).all() and (X <= R).all() and (Y >= -R).all() and (Y <= R).all()
    bits = np.vectorize(lambda v: oracle.mv_bits(4 * int(v)))
    mvc = ((lam * (bits(X) + bits(Y)).astype(np.uint64)) & 0xFFFFFFFF) >> 16
    assert np.array_equal(Cst.astype(np.uint64), S.astype(np.uint64) + mvc)
    hit = (X[:, 592] == 3) & (Y[:, 592] == 2) & (S[:, 592] == 0)
    assert hit.sum() >= 1900
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=os.cpu_count() or 8)
    assert_same((X, Y, S, Cst), want, "4K, all 1980 CTUs")
    pc.free(); pr.free()


@pytest.mark.parametrize("R", [0, 2, 3])
def test_tiny_ranges_and_u8_host_upload(me, oracle, R):
    """Degenerate windows (R = 0: one candidate; a tile narrower than a warp) and the 8-bit host upload entry point."""
    W, H = 192, 64
    f = luma_frames(W, H, 2, seed=31 + R)
    M = R + 8
    cur16, ref16 = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, pad_plane(f[1], M, M, np.uint8)); me.upload(pr, pad_plane(f[0], M, M, np.uint8))
    jobs = frame_jobs(W, H, R, pred=(1, -2))
    me.set_lambda_q16(262144)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur16, (M, M), ref16, (M, M), jobs, R, 262144)
    assert_same(got, want, f"R={R}")
    pc.free(); pr.free()


def test_async_upload_reports_bad_content_at_sync(me):
    W, H, M = 64, 64, 8
    p = me.alloc_plane(1, W, H, M, M)
    bad = np.zeros((H + 2 * M, W + 2 * M), np.int16); bad[5, 5] = -1
    me.upload(p, bad, asynchronous=True)           # enqueue only
    with pytest.raises(hm.HmmeError) as e:
        me.sync()                                  # the deferred 8-bit content check fires here
    assert e.value.code == -5
    me.upload(p, np.zeros_like(bad))               # the context stays usable
    p.free()


def test_randomised_differential_vs_oracle(me, oracle):
    """Seeded fuzz: random ranges (every residue of the tile/round/row-group arithmetic), window offsets, lambdas, content
    classes and plane margins (so the CTU and the window start at arbitrary byte alignments)."""
    g = np.random.default_rng(int(os.environ.get("HMME_FUZZ_SEED", "2026")))
    for it in range(int(os.environ.get("HMME_FUZZ_ITERS", "120"))):       # soak: HMME_FUZZ_ITERS=400 HMME_FUZZ_SEED=n
        R = int(g.integers(0, 45))
        W, H = 64 * int(g.integers(1, 4)), 64 * int(g.integers(1, 3))
        M = R + int(g.integers(6, 23))
        kind = it % 4
        shape = (H + 2 * M, W + 2 * M)
        if kind == 0:
            ref = g.integers(0, 256, size=shape)
        elif kind == 1:                                      # smooth: many near-ties
            ref = (np.add.outer(np.arange(shape[0]) // 3, np.arange(shape[1]) // 5) % 256)
        elif kind == 2:                                      # blocky constant regions: exact ties across tiles
            ref = np.kron(g.integers(0, 256, size=(shape[0] // 16 + 1, shape[1] // 16 + 1)), np.ones((16, 16), int))[:shape[0], :shape[1]]
        else:                                                # binary extremes
            ref = g.integers(0, 2, size=shape) * 255
        ref = ref.astype(np.int16)
        dy, dx = int(g.integers(-3, 4)), int(g.integers(-3, 4))
        cur = np.roll(ref, (dy, dx), (0, 1)).copy()
        cur = np.clip(cur + g.integers(-2, 3, size=shape), 0, 255).astype(np.int16)
        jobs = frame_jobs(W, H, R)
        room = M - R                                         # keep every window inside the allocation
        jobs[:, 2] += g.integers(-room + 1, room, size=len(jobs))
        jobs[:, 3] += g.integers(-room + 1, room, size=len(jobs))
        lam = int(g.choice([0, 1, 65536, 262144, 460000, 4500000, 0x7FFFFFFF]))
        me.set_lambda_q16(lam)
        pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
        me.upload(pc, cur); me.upload(pr, ref)
        got = me.search_frame(pc, pr, jobs, R)
        want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8)
        assert_same(got, want, f"fuzz it={it} R={R} {W}x{H} M={M} kind={kind} lam={lam}")
        pc.free(); pr.free()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_virtual_bands_equal_whole_frame(me, world):
    """band-sharded results must be identical to the single-GPU result.  The same sharding
    code the ranks use (hm.band_jobs / merge_bands) is exercised here with `world` virtual bands on one GPU, each band
    seeing only its own rows of the current frame (the rest of its plane is poisoned)."""
    W, H, R = 448, 320, 12                               # 7 x 5 CTUs: bands cut mid-row
    f = luma_frames(W, H, 2, seed=world)
    M = R + 16
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    me.set_lambda_q16(460000)
    pr, pc = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pr, ref); me.upload(pc, cur)
    whole = me.search_frame(pc, pr, frame_jobs(W, H, R), R)
    parts = []
    for rank in range(world):
        jobs, (r0, r1) = hm.band_jobs(W, H, R, world, rank)
        if len(jobs) == 0:
            parts.append(tuple(np.zeros((0, 593), a.dtype) for a in whole))
            continue
        band = np.full_like(cur, 255)                    # this rank only uploads rows [64*r0, 64*r1) of the current frame
        band[M + 64 * r0:M + 64 * r1] = cur[M + 64 * r0:M + 64 * r1]
        pb = me.alloc_plane(1, W, H, M, M)
        me.upload(pb, band)
        parts.append(me.search_frame(pb, pr, jobs, R))
        pb.free()
    assert_same(hm.merge_bands(parts), whole, f"world={world}")
    pc.free(); pr.free()


def test_cuda_graph_replays_a_whole_step(me, oracle):
    """hmme_graph_*: upload both planes, search, refine, fetch everything recorded once; every replay reads the CURRENT content of
    the page-locked host buffers and must give what the ordinary calls give (here: what the oracle gives)."""
    import torch
    W, H, R, M = 256, 128, 16, 40
    lam = 460000
    me.set_lambda_q16(lam)
    jobs = frame_jobs(W, H, R)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    h_cur, h_ref = pin(np.zeros((H + 2 * M, W + 2 * M), np.int16)), pin(np.zeros((H + 2 * M, W + 2 * M), np.int16))
    outs = [pin(np.zeros((len(jobs), 593), t)) for t in (np.int32, np.int32, np.uint32, np.uint32)]
    fr = pin(np.zeros((len(jobs), 593, 4), np.int32))
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)

    def step():
        me.upload(pr, h_ref, asynchronous=True)
        me.upload(pc, h_cur, asynchronous=True)
        me.search_frame_async(pc, pr, jobs, R)
        me.fetch_results(len(jobs), outs, asynchronous=True)
        me.refine_frame(pc, pr, len(jobs), None, True, asynchronous=True, out=fr)

    step(); me.sync()                                    # sizes every device buffer
    me.graph_begin()
    step()
    g = me.graph_end()
    launches0 = me.kernel_launches
    for seed in (1, 2, 3):
        f = luma_frames(W, H, 2, seed=seed)
        h_cur[:] = pad_plane(f[1], M, M); h_ref[:] = pad_plane(f[0], M, M)
        for o in outs:
            o[:] = 0
        fr[:] = 0
        me.graph_launch(g)
        me.sync()
        want = oracle.search_frame(h_cur, (M, M), h_ref, (M, M), jobs, R, lam, nthreads=4)
        assert_same(outs, want, f"graph replay {seed}")
        rects = me.lib.partition_table()
        pus = np.zeros((len(jobs), 593, 8), np.int32)
        pus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
        pus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
        pus[:, :, 2], pus[:, :, 3] = rects[None, :, 2], rects[None, :, 3]
        pus[:, :, 4], pus[:, :, 5] = want[0], want[1]
        wf = oracle.refine_frac(h_cur, (M, M), h_ref, (M, M), pus.reshape(-1, 8), lam, True)
        assert np.array_equal(fr[:, :, :2].reshape(-1, 2), wf["mvq"]) and np.array_equal(fr[:, :, 2].reshape(-1).view(np.uint32), wf["cost"])
    assert me.kernel_launches == launches0               # replays are not counted as library launches made by the caller
    h_ref[0, 0] = 300                                    # content errors of replayed uploads still surface at the next sync
    me.graph_launch(g)
    with pytest.raises(hm.HmmeError) as e:
        me.sync()
    assert e.value.code == -5
    me.graph_destroy(g)
    with pytest.raises(hm.HmmeError):                    # end without begin
        me.graph_end()
    pc.free(); pr.free()


def test_uploads_are_ordered_after_every_reader_without_a_sync(me, oracle):
    """Asynchronous search -> refinement -> upload of the NEXT frame into the same planes with no hmme_sync in between (the documented
    pipelined sequence): the uploads must wait for the refinement kernel that still reads the planes, and two uploads to one plane
    must keep their order although they alternate over the two io streams."""
    W, H, R, M = 512, 256, 16, 40
    lam = 460000
    me.set_lambda_q16(lam)
    jobs = frame_jobs(W, H, R)
    fr = [luma_frames(W, H, 2, seed=s) for s in (11, 12, 13)]
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    rects = me.lib.partition_table()
    outs, fracs = [], []
    for k, f in enumerate(fr):
        junk = np.full((H + 2 * M, W + 2 * M), 7 * k, np.int16)
        me.upload(pr, junk, asynchronous=True)                   # overwritten by the next upload to the same plane: order must hold
        me.upload(pr, pad_plane(f[0], M, M), asynchronous=True)
        me.upload(pc, pad_plane(f[1], M, M), asynchronous=True)
        me.search_frame_async(pc, pr, jobs, R)
        o = me._outs(len(jobs))
        me.fetch_results(len(jobs), o, asynchronous=True)
        fo = np.zeros((len(jobs), 593, 4), np.int32)
        me.refine_frame(pc, pr, len(jobs), None, True, asynchronous=True, out=fo)
        outs.append(o); fracs.append(fo)
    me.sync()
    for k, f in enumerate(fr):
        cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
        want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=8)
        assert_same(outs[k], want, f"pipelined frame {k}")
        pus = np.zeros((len(jobs), 593, 8), np.int32)
        pus[:, :, 0] = jobs[:, None, 0] + rects[None, :, 0]
        pus[:, :, 1] = jobs[:, None, 1] + rects[None, :, 1]
        pus[:, :, 2], pus[:, :, 3] = rects[None, :, 2], rects[None, :, 3]
        pus[:, :, 4], pus[:, :, 5] = want[0], want[1]
        wf = oracle.refine_frac(cur, (M, M), ref, (M, M), pus.reshape(-1, 8), lam, True)
        assert np.array_equal(fracs[k][:, :, :2].reshape(-1, 2), wf["mvq"]), f"refinement of pipelined frame {k}"
    pc.free(); pr.free()


@pytest.mark.parametrize("R", [0, 3, 32, 33])
def test_window_on_the_last_rows_of_the_allocation(me, oracle, R):
    """Ranges whose candidate-row count is not a multiple of the kernel's 3-row unit, with the window flush against the END of the
    reference allocation (no vertical margin below it): the staging must not fetch the rows that only masked candidates address."""
    W, H = 128, 64
    g = np.random.default_rng(5 + R)
    M = R                                                        # window bottom == last row of the padded plane
    shape = (H + 2 * M, W + 2 * M)
    ref = g.integers(0, 256, size=shape).astype(np.int16)
    cur = np.roll(ref, (1, -2), (0, 1)).copy()
    pc, pr = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M)
    me.upload(pc, cur); me.upload(pr, ref)
    jobs = frame_jobs(W, H, R)                                   # lt = -R: rows [-R, R + 63] = the whole padded height
    me.set_lambda_q16(262144)
    got = me.search_frame(pc, pr, jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, 262144)
    assert_same(got, want, f"flush window R={R}")
    pc.free(); pr.free()


def test_rectangle_upload_equals_whole_plane_upload(me, oracle):
    """hmme_plane_upload_rect_async (int16 and uint8 host planes, odd rectangles): searching with only the band + halo rectangle
    uploaded into a zeroed plane gives what the whole-plane upload gives."""
    W, H, R, M = 384, 256, 9, 32
    f = luma_frames(W, H, 2, seed=77)
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    jobs = frame_jobs(W, H, R)[7:16]                             # a band cut mid-row
    jobs[:, 2] += 3
    me.set_lambda_q16(460000)
    cr, rr = me.lib.band_extent(jobs, R)
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, 460000, nthreads=4)
    for dt in (np.int16, np.uint8):
        pc, pr = me.alloc_plane(1, W, H, 0, 0), me.alloc_plane(1, W, H, M, M)
        me.upload_rect(pr, ref.astype(dt), rr, origin_x=M, origin_y=M)
        me.upload_rect(pc, cur.astype(dt), cr, origin_x=M, origin_y=M)
        got = me.search_frame(pc, pr, jobs, R)
        assert_same(got, want, f"rect upload {dt.__name__}")
        pc.free(); pr.free()
    with pytest.raises(hm.HmmeError):
        me.upload_rect(me.alloc_plane(1, W, H, M, M), ref, (-M - 1, 0, 10, 10), origin_x=M + 1, origin_y=M)


def test_device_resident_tables_keep_both_lists(me, oracle):
    """hmme_table_*: two reference lists and a bi-prediction style 16-bit search of the same frame coexist in one device-resident
    [slot][ctu][593] table (the per-context result buffer would be overwritten by each search)."""
    W, H, R, M = 256, 192, 12, 40
    f = luma_frames(W, H, 3, seed=21)
    cur, r0, r1 = pad_plane(f[1], M, M), pad_plane(f[0], M, M), pad_plane(f[2], M, M)
    bi = (2 * cur.astype(np.int32) - r0).astype(np.int16)
    lam = 460000
    me.set_lambda_q16(lam)
    pc, p0, p1, pb = me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M), me.alloc_plane(1, W, H, M, M), me.alloc_plane(2, W, H, M, M)
    for p, a in ((pc, cur), (p0, r0), (p1, r1), (pb, bi)):
        me.upload(p, a)
    jobs = frame_jobs(W, H, R)
    bij = frame_jobs(W, H, 4, pred=(-3, -2))
    t = me.create_table(3, len(jobs))
    me.search_frame_table(pc, p0, jobs, R, t, 0)
    me.search_frame_table(pc, p1, jobs, R, t, 1)
    me.search_frame_table(pb, p1, bij, 4, t, 2)
    want = [oracle.search_frame(cur, (M, M), r0, (M, M), jobs, R, lam), oracle.search_frame(cur, (M, M), r1, (M, M), jobs, R, lam),
            oracle.search_frame(bi, (M, M), r1, (M, M), bij, 4, lam)]
    for slot in (2, 0, 1):
        assert_same(me.table_fetch(t, slot, 0, len(jobs)), want[slot], f"table slot {slot}")
    part = me.table_fetch(t, 1, 3, 5)
    assert_same(part, [w[3:8] for w in want[1]], "table job range")
    with pytest.raises(hm.HmmeError):
        me.table_fetch(t, 1, 0, len(jobs) + 1)
    me.destroy_table(t)
    for p in (pc, p0, p1, pb):
        p.free()


def test_graphs_survive_plain_calls_and_refuse_stale_buffers(oracle):
    """Advisor findings on hmme_graph_*: (a) after hmme_graph_end, plain asynchronous calls and the timing getters must work although
    events were recorded while capturing; (b) a graph replays its OWN page-locked copy of the job list, so a later per-CTU call or
    another search cannot change what it runs; (c) once a device buffer the graph references has been reallocated, launching it
    is an error instead of a use-after-free."""
    import torch
    W, H, R, M, lam = 256, 128, 8, 24, 460000
    m = hm.MotionEstimator(0, 16)
    m.set_lambda_q16(lam)
    f = luma_frames(W, H, 2, seed=5)
    cur, ref = pad_plane(f[1], M, M), pad_plane(f[0], M, M)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
    h_cur, h_ref = pin(cur), pin(ref)
    jobs = frame_jobs(W, H, R)
    outs = [pin(np.zeros((len(jobs), 593), t)) for t in (np.int32, np.int32, np.uint32, np.uint32)]
    pc, pr = m.alloc_plane(1, W, H, M, M), m.alloc_plane(1, W, H, M, M)

    def step(j):
        m.upload(pr, h_ref, asynchronous=True)
        m.upload(pc, h_cur, asynchronous=True)
        m.search_frame_async(pc, pr, j, R)
        m.fetch_results(len(j), outs, asynchronous=True)

    step(jobs); m.sync()
    m.graph_begin(); step(jobs); g = m.graph_end()
    m.upload(pc, h_cur, asynchronous=True)                      # (a) waits on an event that must not belong to the capture
    m.sync()
    with pytest.raises(hm.HmmeError):
        m.last_kernel_ms()                                       # timing events of the capture are not readable: a clean error, not a CUDA fault
    want = oracle.search_frame(cur, (M, M), ref, (M, M), jobs, R, lam, nthreads=4)
    other = jobs.copy(); other[:, 2] += 3                        # (b) different jobs through the ordinary calls and the per-CTU call in between
    m.search_frame(pc, pr, other, R)
    m.search_ctu(cur[M:M + 64, M:M + 64], ref, 0, 0, M, M, R, -R + 1, -R)
    for o in outs:
        o[:] = 0
    m.graph_launch(g); m.sync()
    assert_same(outs, want, "graph after foreign jobs")
    many = np.tile(jobs, (12, 1))                                # (c) 96 jobs > the context's initial capacity of 64: buffers are reallocated
    m.search_frame(pc, pr, many, R)
    with pytest.raises(hm.HmmeError) as e:
        m.graph_launch(g)
    assert e.value.code == -1 and "reallocated" in str(e.value)
    m.graph_destroy(g)
    pc.free(); pr.free(); m.close()
