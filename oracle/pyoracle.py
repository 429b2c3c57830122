"""ctypes access to the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/hmme_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (hm-opencl_b200/) never does.

  Oracle      : oracle/libhmme_oracle.so   (plain-C restatement; built by `make -C oracle`)
  RefEmu      : oracle/_ref/libhm_ref_me.so (the reference's own TEncOpenCL.cpp + cl/sad.cl over a
                lock-step OpenCL emulation; built by `make -C oracle ref` where /root/reference exists)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_PARTS = 593
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile the checker (gcc only).  `ref` is attempted only where the reference tree exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference/cl"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        # reference-encoder variants (bitstream-parity goldens, B200 drop-in build, CPU-ME baseline); incremental
        subprocess.check_call(["make", "-s", "-C", HERE, "encoders"], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libhmme_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.hmme_oracle_mv_bits.restype = C.c_uint32
        L.hmme_oracle_mv_bits.argtypes = [C.c_int]
        L.hmme_oracle_lambda_q16.restype = C.c_uint32
        L.hmme_oracle_lambda_q16.argtypes = [C.c_double]
        L.hmme_oracle_search_ctu.restype = C.c_int
        L.hmme_oracle_search_ctu.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_uint32] + [C.c_void_p] * 4
        L.hmme_oracle_search_frame.restype = C.c_int
        L.hmme_oracle_search_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_uint32, C.c_int] + [C.c_void_p] * 4
        L.hmme_oracle_mc_cost.restype = C.c_int
        L.hmme_oracle_mc_cost.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hmme_oracle_mc_cost_bi.restype = C.c_int
        L.hmme_oracle_mc_cost_bi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hmme_oracle_refine_frac.restype = C.c_int
        L.hmme_oracle_refine_frac.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint32,
                                              C.c_int] + [C.c_void_p] * 6

    def partition_table(self):
        out = np.zeros((NUM_PARTS, 4), np.int32)
        self.lib.hmme_oracle_partition_table(out.ctypes.data_as(C.c_void_p))
        return out  # rows: x, y, w, h

    def mv_bits(self, v):
        return int(self.lib.hmme_oracle_mv_bits(int(v)))

    def lambda_q16(self, lam):
        return int(self.lib.hmme_oracle_lambda_q16(float(lam)))

    @staticmethod
    def _outs(n):
        return (np.zeros((n, NUM_PARTS), np.int32), np.zeros((n, NUM_PARTS), np.int32),
                np.zeros((n, NUM_PARTS), np.uint32), np.zeros((n, NUM_PARTS), np.uint32))

    def search_ctu(self, cur, plane, ctu_x, ctu_y, origin_x, origin_y, rng, ltx, lty, lam):
        """cur: (64,64) int16.  plane: padded int16 plane whose picture sample (0,0) sits at
        [origin_y, origin_x]; the CTU is at picture position (ctu_x, ctu_y)."""
        cur = np.ascontiguousarray(cur, np.int16)
        assert plane.dtype == np.int16 and plane.flags.c_contiguous
        X, Y, S, Cst = self._outs(1)
        stride = plane.shape[1]
        off = int(((origin_y + ctu_y) * stride + origin_x + ctu_x) * 2)
        rc = self.lib.hmme_oracle_search_ctu(cur.ctypes.data, 64, plane.ctypes.data + off, stride, int(rng), int(ltx), int(lty),
                                             C.c_uint32(lam), X.ctypes.data, Y.ctypes.data, S.ctypes.data, Cst.ctypes.data)
        assert rc == 0
        return X[0], Y[0], S[0], Cst[0]

    def search_frame(self, cur_plane, cur_origin, ref_plane, ref_origin, jobs, rng, lam, nthreads=1):
        """cur_plane/ref_plane int16 2-D; *_origin = (ox, oy) of picture sample (0,0); jobs (n,4) int32
        rows {ctuX, ctuY, ltx, lty}."""
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        n = jobs.shape[0]
        X, Y, S, Cst = self._outs(n)
        cs, rs = cur_plane.shape[1], ref_plane.shape[1]
        co = int((cur_origin[1] * cs + cur_origin[0]) * 2)
        ro = int((ref_origin[1] * rs + ref_origin[0]) * 2)
        rc = self.lib.hmme_oracle_search_frame(cur_plane.ctypes.data + co, cs, ref_plane.ctypes.data + ro, rs,
                                               jobs.ctypes.data, n, rng, C.c_uint32(lam), nthreads,
                                               X.ctypes.data, Y.ctypes.data, S.ctypes.data, Cst.ctypes.data)
        assert rc == 0
        return X, Y, S, Cst


    def refine_frac(self, cur_plane, cur_origin, ref_plane, ref_origin, pus, lam, use_had=True):
        """Fractional-pel refinement (xPatternSearchFracDIF) of a PU list.  Planes int16 2-D, *_origin = (ox, oy) of
        picture sample (0,0); pus (n,8) int32 rows {x, y, w, h, mvx, mvy (integer pel), predx, predy (quarter pel)}.
        Returns dict of mvq (n,2), half (n,2), qter (n,2), cost (n,), dist (n,), cand (n,18) = cost of every candidate."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 8)
        n = pus.shape[0]
        assert cur_plane.dtype == np.int16 and ref_plane.dtype == np.int16
        assert cur_plane.flags.c_contiguous and ref_plane.flags.c_contiguous
        out = dict(mvq=np.zeros((n, 2), np.int32), half=np.zeros((n, 2), np.int32), qter=np.zeros((n, 2), np.int32),
                   cost=np.zeros(n, np.uint32), dist=np.zeros(n, np.uint32), cand=np.zeros((n, 18), np.uint32))
        cs, rs = cur_plane.shape[1], ref_plane.shape[1]
        co = int((cur_origin[1] * cs + cur_origin[0]) * 2)
        ro = int((ref_origin[1] * rs + ref_origin[0]) * 2)
        rc = self.lib.hmme_oracle_refine_frac(cur_plane.ctypes.data + co, cs, ref_plane.ctypes.data + ro, rs, pus.ctypes.data, n,
                                              C.c_uint32(lam), int(bool(use_had)), out["mvq"].ctypes.data, out["half"].ctypes.data,
                                              out["qter"].ctypes.data, out["cost"].ctypes.data, out["dist"].ctypes.data, out["cand"].ctypes.data)
        assert rc == 0
        return out


    def mc_cost(self, cur_plane, cur_origin, ref_plane, ref_origin, pus, use_had=False):
        """Distortion of the motion-compensated uni-prediction; pus (n,6) int32 rows {x, y, w, h, mvx, mvy (quarter pel)}."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 6)
        assert cur_plane.dtype == np.int16 and ref_plane.dtype == np.int16 and cur_plane.flags.c_contiguous and ref_plane.flags.c_contiguous
        out = np.zeros(pus.shape[0], np.uint32)
        cs, rs = cur_plane.shape[1], ref_plane.shape[1]
        rc = self.lib.hmme_oracle_mc_cost(cur_plane.ctypes.data + int((cur_origin[1] * cs + cur_origin[0]) * 2), cs,
                                          ref_plane.ctypes.data + int((ref_origin[1] * rs + ref_origin[0]) * 2), rs,
                                          pus.ctypes.data, pus.shape[0], int(bool(use_had)), out.ctypes.data)
        assert rc == 0
        return out


    def mc_cost_bi(self, cur_plane, cur_origin, ref0_plane, ref1_plane, ref_origin, pus, use_had=False):
        """Bi-directional form; pus (n,8) int32 rows {x, y, w, h, mv0x, mv0y, mv1x, mv1y}; both reference planes share ref_origin."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 8)
        for a in (cur_plane, ref0_plane, ref1_plane):
            assert a.dtype == np.int16 and a.flags.c_contiguous
        out = np.zeros(pus.shape[0], np.uint32)
        cs, r0s, r1s = cur_plane.shape[1], ref0_plane.shape[1], ref1_plane.shape[1]
        rc = self.lib.hmme_oracle_mc_cost_bi(cur_plane.ctypes.data + int((cur_origin[1] * cs + cur_origin[0]) * 2), cs,
                                             ref0_plane.ctypes.data + int((ref_origin[1] * r0s + ref_origin[0]) * 2), r0s,
                                             ref1_plane.ctypes.data + int((ref_origin[1] * r1s + ref_origin[0]) * 2), r1s,
                                             pus.ctypes.data, pus.shape[0], int(bool(use_had)), out.ctypes.data)
        assert rc == 0
        return out


class RefEmu:
    """The reference's own host class + kernels, lock-step on the CPU.  available() is False on the
    GPU box (no /root/reference there and hence possibly no oracle/_ref)."""
    PATH = os.path.join(HERE, "_ref", "libhm_ref_me.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self, kernel_path, search_range):
        L = self.lib = C.CDLL(self.PATH)
        L.hmref_create.restype = C.c_void_p
        L.hmref_create.argtypes = [C.c_char_p, C.c_int]
        L.hmref_destroy.argtypes = [C.c_void_p]
        L.hmref_set_lambda.argtypes = [C.c_void_p, C.c_double]
        L.hmref_set_lambda_q16.argtypes = [C.c_void_p, C.c_uint32]
        L.hmref_get_lambda_q16.restype = C.c_uint32
        L.hmref_get_lambda_q16.argtypes = [C.c_void_p]
        L.hmref_calc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.refemu_stats.argtypes = [C.POINTER(C.c_long)] * 3
        self.h = L.hmref_create(kernel_path.encode(), search_range)
        if not self.h:
            raise RuntimeError("reference TEncOpenCL init failed")

    def close(self):
        if self.h:
            self.lib.hmref_destroy(self.h)
            self.h = None

    def set_lambda(self, lam):
        self.lib.hmref_set_lambda(self.h, float(lam))
        return int(self.lib.hmref_get_lambda_q16(self.h))

    def set_lambda_q16(self, v):
        self.lib.hmref_set_lambda_q16(self.h, C.c_uint32(v))

    def calc(self, cur, plane, ctu_x, ctu_y, origin_x, origin_y, rng, ltx, lty):
        cur = np.ascontiguousarray(cur, np.int16)
        X = np.zeros(NUM_PARTS, np.int32); Y = np.zeros(NUM_PARTS, np.int32)
        S = np.zeros(NUM_PARTS, np.uint32); Cst = np.zeros(NUM_PARTS, np.uint32)
        stride = plane.shape[1]
        off = int(((origin_y + ctu_y) * stride + origin_x + ctu_x) * 2)
        self.lib.hmref_calc(self.h, cur.ctypes.data, plane.ctypes.data + off, stride, rng, ltx, lty,
                            X.ctypes.data, Y.ctypes.data, S.ctypes.data, Cst.ctypes.data)
        return X, Y, S, Cst

    def stats(self):
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        self.lib.refemu_stats(C.byref(a), C.byref(b), C.byref(c))
        return {"oob_reads": a.value, "oob_writes": b.value, "launches": c.value}
