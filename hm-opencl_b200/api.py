"""ctypes binding of libhmme_b200.so + the Python mirror of the reference's TEncOpenCL surface.

Reference interface mirrored (names, argument meaning, error behaviour):
  /root/reference/source/Lib/TLibEncoder/TEncOpenCL.h:105-123
      findDevice / compileKernelSource / createBuffers / calcMotionVectors /
      getX / getY / getRuiCost / setLambda / setEnabled / getDeviceInfo
The compiled drop-in for the encoder itself is the C++ class in hm-opencl_b200/host/.
"""
import ctypes as C
import os

import numpy as np

NUM_CTU_PARTS = 593
_HERE = os.path.dirname(os.path.abspath(__file__))

# every symbol include/hmme_b200.h declares (tests check header == this list == the .so's exports)
EXPORTS = [
    "hmme_device_count", "hmme_create", "hmme_destroy", "hmme_device_name", "hmme_last_error", "hmme_stream",
    "hmme_set_lambda", "hmme_set_lambda_q16", "hmme_get_lambda_q16", "hmme_search_ctu",
    "hmme_plane_alloc", "hmme_plane_free", "hmme_plane_upload_s16", "hmme_plane_upload_u8",
    "hmme_search_frame", "hmme_search_frame_async", "hmme_fetch_results", "hmme_sync",
    "hmme_plane_upload_s16_async", "hmme_fetch_results_async",
    "hmme_refine_pu", "hmme_refine_frac", "hmme_refine_frame", "hmme_refine_frame_async", "hmme_fetch_frac_async", "hmme_last_frac_ms", "hmme_mc_cost", "hmme_mc_cost_pu", "hmme_mc_cost_bi", "hmme_mc_cost_bi_pu",
    "hmme_graph_begin", "hmme_graph_end", "hmme_graph_launch", "hmme_graph_destroy",
    "hmme_plane_upload_u8_async", "hmme_plane_upload_rect_async", "hmme_host_alloc", "hmme_host_free",
    "hmme_table_create", "hmme_table_destroy", "hmme_search_frame_table_async", "hmme_table_fetch_async", "hmme_table_device_ptr",
    "hmme_group_create", "hmme_group_unique_id", "hmme_group_create_rank", "hmme_group_destroy", "hmme_group_last_error", "hmme_group_size",
    "hmme_group_set_lambda_q16", "hmme_group_configure", "hmme_group_search_frame_async", "hmme_group_sync", "hmme_group_search_frame",
    "hmme_group_band", "hmme_group_pipeline_depth", "hmme_search_launch_size", "hmme_group_context", "hmme_group_last_kernel_ms", "hmme_group_kernel_launches", "hmme_band_split", "hmme_band_extent",
    "hmme_last_kernel_ms", "hmme_kernel_launches", "hmme_measure_int_alu_peak", "hmme_partition_rect", "hmme_index_block", "hmme_search_window", "hmme_version",
]


class HmmeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hmme error %d: %s" % (code, msg))
        self.code = code


class PlaneDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("elemBytes", C.c_int32), ("pitch", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("marginX", C.c_int32), ("marginY", C.c_int32)]


def lib_path():
    # HMME_B200_LIB: another build of the same library (kernel experiments); never a different implementation
    return os.environ.get("HMME_B200_LIB") or os.path.join(_HERE, "libhmme_b200.so")


class HmmeLib:
    """The loaded shared library.  Missing library == hard error (there is no fallback path)."""
    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        path = lib_path()
        if not os.path.exists(path):
            raise HmmeError(-100, "CUDA extension %s is missing: run `python hm-opencl_b200/build.py` "
                                  "(this package has no CPU fallback)" % path)
        L = self.L = C.CDLL(path)
        vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
        P = C.POINTER
        sig = {
            "hmme_device_count": (i32, [P(C.c_int)]),
            "hmme_create": (i32, [P(vp), i32, i32, i32, i32]),
            "hmme_destroy": (None, [vp]),
            "hmme_device_name": (C.c_char_p, [vp]),
            "hmme_last_error": (C.c_char_p, [vp]),
            "hmme_stream": (vp, [vp]),
            "hmme_set_lambda": (i32, [vp, C.c_double]),
            "hmme_set_lambda_q16": (i32, [vp, u32]),
            "hmme_get_lambda_q16": (u32, [vp]),
            "hmme_search_ctu": (i32, [vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
            "hmme_plane_alloc": (i32, [vp, P(PlaneDesc), i32, i32, i32, i32, i32]),
            "hmme_plane_free": (i32, [vp, P(PlaneDesc)]),
            "hmme_plane_upload_s16": (i32, [vp, P(PlaneDesc), vp, i32]),
            "hmme_plane_upload_u8": (i32, [vp, P(PlaneDesc), vp, i32]),
            "hmme_search_frame": (i32, [vp, P(PlaneDesc), P(PlaneDesc), vp, i32, i32, vp, vp, vp, vp]),
            "hmme_search_frame_async": (i32, [vp, P(PlaneDesc), P(PlaneDesc), vp, i32, i32]),
            "hmme_fetch_results": (i32, [vp, i32, vp, vp, vp, vp]),
            "hmme_fetch_results_async": (i32, [vp, i32, vp, vp, vp, vp]),
            "hmme_plane_upload_s16_async": (i32, [vp, P(PlaneDesc), vp, i32]),
            "hmme_sync": (i32, [vp]),
            "hmme_refine_pu": (i32, [vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, P(C.c_int32), P(C.c_int32), P(u32), P(u32)]),
            "hmme_refine_frac": (i32, [vp, P(PlaneDesc), P(PlaneDesc), vp, i32, i32, vp, vp]),
            "hmme_refine_frame": (i32, [vp, P(PlaneDesc), P(PlaneDesc), i32, vp, i32, vp]),
            "hmme_refine_frame_async": (i32, [vp, P(PlaneDesc), P(PlaneDesc), i32, vp, i32]),
            "hmme_fetch_frac_async": (i32, [vp, i32, vp]),
            "hmme_last_frac_ms": (i32, [vp, P(C.c_float)]),
            "hmme_graph_begin": (i32, [vp]),
            "hmme_graph_end": (i32, [vp, P(vp)]),
            "hmme_graph_launch": (i32, [vp, vp]),
            "hmme_graph_destroy": (None, [vp]),
            "hmme_mc_cost_pu": (i32, [vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, P(u32)]),
            "hmme_mc_cost_bi": (i32, [vp, P(PlaneDesc), P(PlaneDesc), P(PlaneDesc), vp, i32, i32, vp]),
            "hmme_mc_cost_bi_pu": (i32, [vp, vp, i32, vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, P(u32)]),
            "hmme_mc_cost": (i32, [vp, P(PlaneDesc), P(PlaneDesc), vp, i32, i32, vp]),
            "hmme_host_alloc": (vp, [C.c_size_t]),
            "hmme_host_free": (None, [vp]),
            "hmme_plane_upload_u8_async": (i32, [vp, P(PlaneDesc), vp, i32]),
            "hmme_plane_upload_rect_async": (i32, [vp, P(PlaneDesc), vp, i32, i32, i32, i32, i32, i32]),
            "hmme_table_create": (i32, [vp, P(vp), i32, i32]),
            "hmme_table_destroy": (None, [vp]),
            "hmme_search_frame_table_async": (i32, [vp, P(PlaneDesc), P(PlaneDesc), vp, i32, i32, vp, i32]),
            "hmme_table_fetch_async": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp]),
            "hmme_table_device_ptr": (vp, [vp, i32, i32]),
            "hmme_group_create": (i32, [P(vp), P(C.c_int), i32, i32]),
            "hmme_group_unique_id": (i32, [vp, C.c_size_t]),
            "hmme_group_create_rank": (i32, [P(vp), i32, i32, i32, vp, i32]),
            "hmme_group_destroy": (None, [vp]),
            "hmme_group_last_error": (C.c_char_p, [vp]),
            "hmme_group_size": (i32, [vp, P(C.c_int), P(C.c_int)]),
            "hmme_group_set_lambda_q16": (i32, [vp, u32]),
            "hmme_group_configure": (i32, [vp, i32, i32, i32, i32, i32]),
            "hmme_group_search_frame_async": (i32, [vp, i32, vp, i32, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp]),
            "hmme_group_sync": (i32, [vp, i32]),
            "hmme_group_search_frame": (i32, [vp, vp, i32, vp, i32, i32, vp, i32, i32, vp, vp, vp, vp]),
            "hmme_group_band": (i32, [vp, i32, i32, P(C.c_int), P(C.c_int)]),
            "hmme_group_pipeline_depth": (i32, [vp, i32, i32]),
            "hmme_search_launch_size": (i32, [vp, i32, i32, P(C.c_int), P(C.c_int)]),
            "hmme_group_context": (vp, [vp, i32, i32]),
            "hmme_group_last_kernel_ms": (i32, [vp, i32, P(C.c_float)]),
            "hmme_group_kernel_launches": (C.c_uint64, [vp]),
            "hmme_band_split": (i32, [i32, i32, i32, P(C.c_int), P(C.c_int)]),
            "hmme_band_extent": (i32, [vp, i32, i32, vp, vp]),
            "hmme_last_kernel_ms": (i32, [vp, P(C.c_float)]),
            "hmme_kernel_launches": (C.c_uint64, [vp]),
            "hmme_measure_int_alu_peak": (i32, [vp, P(C.c_double), P(C.c_double), P(C.c_double)]),
            "hmme_partition_rect": (i32, [i32, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int)]),
            "hmme_index_block": (i32, [i32] * 6),
            "hmme_search_window": (i32, [i32] * 7 + [P(C.c_int)] * 4),
            "hmme_version": (C.c_char_p, []),
        }
        assert sorted(sig) == sorted(EXPORTS)
        for name, (res, args) in sig.items():
            fn = getattr(L, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args

    def partition_table(self):
        out = np.zeros((NUM_CTU_PARTS, 4), np.int32)
        v = [C.c_int() for _ in range(4)]
        for p in range(NUM_CTU_PARTS):
            assert self.L.hmme_partition_rect(p, *[C.byref(q) for q in v]) == 0
            out[p] = [q.value for q in v]
        return out

    def index_block(self, part_size, depth, part_idx, z_idx, cu_w, cu_h):
        return int(self.L.hmme_index_block(part_size, depth, part_idx, z_idx, cu_w, cu_h))

    def search_window(self, pred_hor_qpel, pred_ver_qpel, rng, cu_x, cu_y, pic_w, pic_h):
        v = [C.c_int() for _ in range(4)]
        assert self.L.hmme_search_window(pred_hor_qpel, pred_ver_qpel, rng, cu_x, cu_y, pic_w, pic_h, *[C.byref(q) for q in v]) == 0
        return tuple(q.value for q in v)      # ltx, lty, rbx, rby

    def band_split(self, njobs, world, rank):
        """Jobs [first, first + count) of njobs belong to `rank` of `world` (hmme_band_split: the split the group uses)."""
        f, n = C.c_int(), C.c_int()
        rc = self.L.hmme_band_split(int(njobs), int(world), int(rank), C.byref(f), C.byref(n))
        if rc != 0:
            raise HmmeError(rc, "hmme_band_split: bad argument")
        return f.value, n.value

    def band_extent(self, jobs, rng):
        """Picture rectangles (x0, y0, x1, y1) a job range reads: its CTUs of the current frame, band + halo of the reference."""
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        cr, rr = np.zeros(4, np.int32), np.zeros(4, np.int32)
        rc = self.L.hmme_band_extent(jobs.ctypes.data, jobs.shape[0], int(rng), cr.ctypes.data, rr.ctypes.data)
        if rc != 0:
            raise HmmeError(rc, "hmme_band_extent: bad argument")
        return tuple(int(v) for v in cr), tuple(int(v) for v in rr)

    def device_count(self):
        n = C.c_int(0)
        rc = self.L.hmme_device_count(C.byref(n))
        if rc != 0:
            raise HmmeError(rc, self.L.hmme_last_error(None).decode())
        return n.value


class Plane:
    """A device-resident luma plane (library-owned, or a view of external device memory such as a torch tensor)."""

    def __init__(self, me, desc, owned):
        self.me, self.desc, self.owned = me, desc, owned

    @property
    def nbytes(self):
        d = self.desc
        return d.pitch * (d.height + 2 * d.marginY) * d.elemBytes

    def free(self):
        if self.owned and self.desc.base:
            self.me._chk(self.me.lib.L.hmme_plane_free(self.me.h, C.byref(self.desc)))
        self.desc.base = None


class MotionEstimator:
    """One hmme context (one GPU, one stream)."""

    def __init__(self, device=0, max_search_range=64):
        self.lib = HmmeLib.get()
        h = C.c_void_p()
        rc = self.lib.L.hmme_create(C.byref(h), device, 64, 64, max_search_range)
        if rc != 0:
            raise HmmeError(rc, self.lib.L.hmme_last_error(None).decode())
        self.h = h
        self.max_search_range = max_search_range

    @classmethod
    def borrowed(cls, handle, owner):
        """Wrapper of a context somebody else owns (hmme_group_context): close() leaves it alone."""
        me = cls.__new__(cls)
        me.lib, me.h, me.max_search_range, me._owner = HmmeLib.get(), C.c_void_p(handle), 0, owner
        return me

    def close(self):
        if getattr(self, "h", None) and getattr(self, "_owner", None) is None:
            self.lib.L.hmme_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise HmmeError(rc, self.lib.L.hmme_last_error(self.h).decode())

    # -- properties
    @property
    def device_name(self):
        return self.lib.L.hmme_device_name(self.h).decode()

    @property
    def stream_ptr(self):
        return int(self.lib.L.hmme_stream(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.L.hmme_kernel_launches(self.h))

    def set_lambda(self, lam):
        self._chk(self.lib.L.hmme_set_lambda(self.h, float(lam)))

    def set_lambda_q16(self, v):
        self._chk(self.lib.L.hmme_set_lambda_q16(self.h, C.c_uint32(int(v))))

    def get_lambda_q16(self):
        return int(self.lib.L.hmme_get_lambda_q16(self.h))

    # -- synchronous per-CTU search (TEncOpenCL::calcMotionVectors)
    def search_ctu(self, cur, plane, ctu_x, ctu_y, origin_x, origin_y, rng, ltx, lty):
        cur = np.ascontiguousarray(cur, np.int16)
        assert cur.shape == (64, 64) and plane.dtype == np.int16 and plane.flags.c_contiguous
        out = [np.zeros(NUM_CTU_PARTS, t) for t in (np.int32, np.int32, np.uint32, np.uint32)]
        stride = plane.shape[1]
        off = int(((origin_y + ctu_y) * stride + origin_x + ctu_x) * 2)
        self._chk(self.lib.L.hmme_search_ctu(self.h, cur.ctypes.data, 64, plane.ctypes.data + off, stride, int(rng), int(ltx),
                                             int(lty), *[o.ctypes.data for o in out]))
        return tuple(out)

    # -- planes
    def alloc_plane(self, elem_bytes, width, height, margin_x, margin_y):
        d = PlaneDesc()
        self._chk(self.lib.L.hmme_plane_alloc(self.h, C.byref(d), elem_bytes, width, height, margin_x, margin_y))
        return Plane(self, d, True)

    def wrap_plane(self, device_ptr, elem_bytes, pitch, width, height, margin_x, margin_y):
        """View of device memory owned by someone else (e.g. torch.Tensor.data_ptr()); must be 16-byte aligned
        and hold pitch*(height+2*margin_y) elements plus 64 bytes of slack (16-byte granular TMA row copies)."""
        d = PlaneDesc(C.c_void_p(int(device_ptr)), elem_bytes, pitch, width, height, margin_x, margin_y)
        return Plane(self, d, False)

    def upload(self, plane, host, origin_x=None, origin_y=None, asynchronous=False):
        """host: 2-D numpy array (int16 or uint8) that includes the margins; picture sample (0,0) at
        [origin_y, origin_x] (defaults: the plane's margins)."""
        d = plane.desc
        ox = d.marginX if origin_x is None else origin_x
        oy = d.marginY if origin_y is None else origin_y
        assert host.ndim == 2 and host.flags.c_contiguous
        assert oy >= d.marginY and ox >= d.marginX and host.shape[0] - oy >= d.height + d.marginY and host.shape[1] - ox >= d.width + d.marginX
        off = int(oy * host.shape[1] + ox) * host.itemsize
        if host.dtype == np.int16:
            fn = self.lib.L.hmme_plane_upload_s16_async if asynchronous else self.lib.L.hmme_plane_upload_s16
            self._chk(fn(self.h, C.byref(d), host.ctypes.data + off, host.shape[1]))
        elif host.dtype == np.uint8:
            fn = self.lib.L.hmme_plane_upload_u8_async if asynchronous else self.lib.L.hmme_plane_upload_u8
            self._chk(fn(self.h, C.byref(d), host.ctypes.data + off, host.shape[1]))
        else:
            raise TypeError("plane uploads take int16 (HM Pel) or uint8 arrays")

    def upload_rect(self, plane, host, rect, origin_x=None, origin_y=None):
        """Asynchronous upload of the picture rectangle rect = (x0, y0, x1, y1) only (hmme_plane_upload_rect_async)."""
        d = plane.desc
        ox = d.marginX if origin_x is None else origin_x
        oy = d.marginY if origin_y is None else origin_y
        assert host.ndim == 2 and host.flags.c_contiguous and host.dtype in (np.int16, np.uint8)
        x0, y0, x1, y1 = (int(v) for v in rect)
        assert oy + y0 >= 0 and ox + x0 >= 0 and oy + y1 <= host.shape[0] and ox + x1 <= host.shape[1]
        off = int(oy * host.shape[1] + ox) * host.itemsize
        self._chk(self.lib.L.hmme_plane_upload_rect_async(self.h, C.byref(d), host.ctypes.data + off, host.shape[1], host.itemsize, x0, y0, x1, y1))

    # -- device-resident result tables (allMotionVectors / allRuiCost for every CTU of a picture, one slot per list / reference / hypothesis)
    def create_table(self, slots, jobs_per_slot):
        t = C.c_void_p()
        self._chk(self.lib.L.hmme_table_create(self.h, C.byref(t), int(slots), int(jobs_per_slot)))
        return t

    def destroy_table(self, t):
        self.lib.L.hmme_table_destroy(t)

    def search_frame_table(self, cur, ref, jobs, rng, table, slot):
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        self._chk(self.lib.L.hmme_search_frame_table_async(self.h, C.byref(cur.desc), C.byref(ref.desc), jobs.ctypes.data, jobs.shape[0], int(rng), table, int(slot)))
        self._keep_jobs = jobs

    def table_fetch(self, table, slot, first, njobs, out=None):
        out = out or self._outs(njobs)
        self._chk(self.lib.L.hmme_table_fetch_async(self.h, table, int(slot), int(first), int(njobs), *[o.ctypes.data for o in out]))
        self.sync()
        return tuple(out)

    # -- whole-frame batch
    @staticmethod
    def _outs(n):
        return [np.zeros((n, NUM_CTU_PARTS), t) for t in (np.int32, np.int32, np.uint32, np.uint32)]

    def search_frame(self, cur, ref, jobs, rng):
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        out = self._outs(jobs.shape[0])
        self._chk(self.lib.L.hmme_search_frame(self.h, C.byref(cur.desc), C.byref(ref.desc), jobs.ctypes.data, jobs.shape[0],
                                               int(rng), *[o.ctypes.data for o in out]))
        return tuple(out)

    def search_frame_async(self, cur, ref, jobs, rng):
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        self._chk(self.lib.L.hmme_search_frame_async(self.h, C.byref(cur.desc), C.byref(ref.desc), jobs.ctypes.data,
                                                     jobs.shape[0], int(rng)))
        return jobs.shape[0]

    def fetch_results(self, njobs, out=None, asynchronous=False):
        out = out or self._outs(njobs)
        fn = self.lib.L.hmme_fetch_results_async if asynchronous else self.lib.L.hmme_fetch_results
        self._chk(fn(self.h, njobs, *[o.ctypes.data for o in out]))
        return tuple(out)

    def sync(self):
        self._chk(self.lib.L.hmme_sync(self.h))

    # -- pre-marshalled asynchronous calls for launch-bound loops: the same C entry points, their ctypes arguments built once
    def _bound(self, fn, *args):
        chk = self._chk

        def call():
            chk(fn(*args))
        call.keepalive = args
        return call

    def bind_upload(self, plane, host, origin_x=None, origin_y=None):
        """Zero-argument callable = upload(plane, host, ..., asynchronous=True); `host` (int16, page-locked) must stay alive and in place."""
        d = plane.desc
        ox = d.marginX if origin_x is None else origin_x
        oy = d.marginY if origin_y is None else origin_y
        assert host.dtype == np.int16 and host.ndim == 2 and host.flags.c_contiguous
        assert oy >= d.marginY and ox >= d.marginX and host.shape[0] - oy >= d.height + d.marginY and host.shape[1] - ox >= d.width + d.marginX
        off = int(oy * host.shape[1] + ox) * host.itemsize
        f = self._bound(self.lib.L.hmme_plane_upload_s16_async, self.h, C.byref(d), C.c_void_p(host.ctypes.data + off), C.c_int(host.shape[1]))
        f.host = host
        return f

    def bind_search(self, cur, ref, jobs, rng):
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        f = self._bound(self.lib.L.hmme_search_frame_async, self.h, C.byref(cur.desc), C.byref(ref.desc), C.c_void_p(jobs.ctypes.data),
                        C.c_int(jobs.shape[0]), C.c_int(int(rng)))
        f.jobs = jobs
        return f

    def bind_fetch(self, njobs, outs):
        f = self._bound(self.lib.L.hmme_fetch_results_async, self.h, C.c_int(njobs), *[C.c_void_p(o.ctypes.data) for o in outs])
        f.outs = outs
        return f

    def bind_sync(self):
        return self._bound(self.lib.L.hmme_sync, self.h)

    # -- CUDA graphs: record the asynchronous calls of a step once, replay them with one call
    def graph_begin(self):
        self._chk(self.lib.L.hmme_graph_begin(self.h))

    def graph_end(self):
        g = C.c_void_p()
        self._chk(self.lib.L.hmme_graph_end(self.h, C.byref(g)))
        return g

    def graph_launch(self, g):
        self._chk(self.lib.L.hmme_graph_launch(self.h, g))

    def graph_destroy(self, g):
        self.lib.L.hmme_graph_destroy(g)

    # -- fractional-pel refinement (TEncSearch::xPatternSearchFracDIF, TEncSearch.cpp:4294-4331)
    FRAC_DTYPE = np.dtype([("mvx", np.int32), ("mvy", np.int32), ("cost", np.uint32), ("dist", np.uint32)])

    def refine_frac(self, cur, ref, pus, use_had=True, want_candidates=False):
        """pus: (n, 8) int32 rows {x, y, w, h, mvx, mvy (integer pel), predx, predy (quarter pel)}.  Returns a structured
        array (mvx, mvy quarter-pel; cost; dist) and, on request, the (n, 18) candidate costs in the reference's table order."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 8)
        n = pus.shape[0]
        res = np.zeros(n, self.FRAC_DTYPE)
        cand = np.zeros((n, 18), np.uint32) if want_candidates else None
        self._chk(self.lib.L.hmme_refine_frac(self.h, C.byref(cur.desc), C.byref(ref.desc), pus.ctypes.data, n, int(bool(use_had)),
                                              res.ctypes.data, cand.ctypes.data if want_candidates else None))
        return (res, cand) if want_candidates else res

    def refine_pu(self, cur_block, ref_plane, pu_x, pu_y, origin_x, origin_y, mv, pred, use_had=True):
        """Host arrays, synchronous (the form xPatternSearchFracDIF has): cur_block (h, w) int16; ref_plane padded int16 plane
        whose picture sample (0,0) sits at [origin_y, origin_x]; the PU is at picture position (pu_x, pu_y).
        Returns (mvx, mvy quarter-pel, cost, dist)."""
        cur_block = np.ascontiguousarray(cur_block, np.int16)
        assert ref_plane.dtype == np.int16 and ref_plane.flags.c_contiguous
        h, w = cur_block.shape
        stride = ref_plane.shape[1]
        off = int(((origin_y + pu_y) * stride + origin_x + pu_x) * 2)
        mx, my, cost, dist = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32()
        self._chk(self.lib.L.hmme_refine_pu(self.h, cur_block.ctypes.data, w, ref_plane.ctypes.data + off, stride, w, h, int(mv[0]), int(mv[1]),
                                            int(pred[0]), int(pred[1]), int(bool(use_had)), C.byref(mx), C.byref(my), C.byref(cost), C.byref(dist)))
        return mx.value, my.value, cost.value, dist.value

    def refine_frame(self, cur, ref, njobs, preds=None, use_had=True, asynchronous=False, out=None):
        """All 593 partitions of every job of the preceding search_frame[_async] on this context, from its integer winners
        (which stay on the device).  preds: None or (njobs, 2) quarter-pel predictors.  Returns (njobs, 593) structured."""
        if preds is not None:
            preds = np.ascontiguousarray(preds, np.int32).reshape(njobs, 2)
        pp = preds.ctypes.data if preds is not None else None
        if asynchronous:
            self._keep = preds                                  # the copy is enqueued from this array: keep it alive until sync()
            self._chk(self.lib.L.hmme_refine_frame_async(self.h, C.byref(cur.desc), C.byref(ref.desc), int(njobs), pp, int(bool(use_had))))
            if out is not None:
                self._chk(self.lib.L.hmme_fetch_frac_async(self.h, int(njobs), out.ctypes.data))
            return out
        res = out if out is not None else np.zeros((njobs, NUM_CTU_PARTS), self.FRAC_DTYPE)
        self._chk(self.lib.L.hmme_refine_frame(self.h, C.byref(cur.desc), C.byref(ref.desc), int(njobs), pp, int(bool(use_had)), res.ctypes.data))
        return res

    def mc_cost(self, cur, ref, pus, use_had=False):
        """Distortion of the motion-compensated uni-prediction (xGetTemplateCost / uni merge candidates): pus (n, 6) int32 rows
        {x, y, w, h, mvx, mvy (QUARTER pel, already clipped)}.  Returns (n,) uint32: SAD, or Hadamard SATD with use_had."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 6)
        out = np.zeros(pus.shape[0], np.uint32)
        self._chk(self.lib.L.hmme_mc_cost(self.h, C.byref(cur.desc), C.byref(ref.desc), pus.ctypes.data, pus.shape[0], int(bool(use_had)), out.ctypes.data))
        return out

    def mc_cost_bi(self, cur, ref0, ref1, pus, use_had=False):
        """Bi-directional form: pus (n, 8) int32 rows {x, y, w, h, mv0x, mv0y, mv1x, mv1y} (quarter pel, clipped)."""
        pus = np.ascontiguousarray(pus, np.int32).reshape(-1, 8)
        out = np.zeros(pus.shape[0], np.uint32)
        self._chk(self.lib.L.hmme_mc_cost_bi(self.h, C.byref(cur.desc), C.byref(ref0.desc), C.byref(ref1.desc), pus.ctypes.data, pus.shape[0],
                                             int(bool(use_had)), out.ctypes.data))
        return out

    def mc_cost_bi_pu(self, cur_block, ref0_plane, ref1_plane, pu_x, pu_y, origin_x, origin_y, mv0, mv1, use_had=False):
        """Host arrays, synchronous, two reference planes of equal geometry."""
        cur_block = np.ascontiguousarray(cur_block, np.int16)
        h, w = cur_block.shape
        s0, s1 = ref0_plane.shape[1], ref1_plane.shape[1]
        o0 = int(((origin_y + pu_y) * s0 + origin_x + pu_x) * 2)
        o1 = int(((origin_y + pu_y) * s1 + origin_x + pu_x) * 2)
        d = C.c_uint32()
        self._chk(self.lib.L.hmme_mc_cost_bi_pu(self.h, cur_block.ctypes.data, w, ref0_plane.ctypes.data + o0, s0, int(mv0[0]), int(mv0[1]),
                                                ref1_plane.ctypes.data + o1, s1, int(mv1[0]), int(mv1[1]), w, h, int(bool(use_had)), C.byref(d)))
        return d.value

    def mc_cost_pu(self, cur_block, ref_plane, pu_x, pu_y, origin_x, origin_y, mv_qpel, use_had=False):
        """Host arrays, synchronous: distortion of the prediction of one PU at a clipped quarter-pel MV (xGetTemplateCost's arguments)."""
        cur_block = np.ascontiguousarray(cur_block, np.int16)
        assert ref_plane.dtype == np.int16 and ref_plane.flags.c_contiguous
        h, w = cur_block.shape
        stride = ref_plane.shape[1]
        off = int(((origin_y + pu_y) * stride + origin_x + pu_x) * 2)
        d = C.c_uint32()
        self._chk(self.lib.L.hmme_mc_cost_pu(self.h, cur_block.ctypes.data, w, ref_plane.ctypes.data + off, stride, w, h, int(mv_qpel[0]), int(mv_qpel[1]),
                                             int(bool(use_had)), C.byref(d)))
        return d.value

    def last_frac_ms(self):
        ms = C.c_float()
        self._chk(self.lib.L.hmme_last_frac_ms(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self._chk(self.lib.L.hmme_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def measure_int_alu_peak(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.lib.L.hmme_measure_int_alu_peak(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"lane_ops_per_s": a.value, "lanes_per_clk_sm": b.value, "sm_mhz": c.value}


class Group:
    """hmme_group: one frame over several GPUs behind the C ABI (band split, band + halo or NCCL-broadcast reference distribution,
    results into one host table).  devices=[...] drives them from this process; rank/world/unique_id makes this process one rank of
    a one-process-per-GPU job (torchrun)."""
    BAND_HALO, BROADCAST = 0, 1
    SLOTS = 3                  # HMME_GROUP_SLOTS: frames in flight per group

    def __init__(self, devices=None, max_search_range=64, device=None, rank=None, world=None, unique_id=None):
        self.lib = HmmeLib.get()
        g = C.c_void_p()
        if rank is None:
            devs = (C.c_int * len(devices))(*devices)
            rc = self.lib.L.hmme_group_create(C.byref(g), devs, len(devices), max_search_range)
        else:
            self._uid = (C.c_char * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
            rc = self.lib.L.hmme_group_create_rank(C.byref(g), int(device), int(rank), int(world), self._uid, max_search_range)
        if rc != 0:
            raise HmmeError(rc, self.lib.L.hmme_group_last_error(None).decode())
        self.g = g
        w, n = C.c_int(), C.c_int()
        self.lib.L.hmme_group_size(g, C.byref(w), C.byref(n))
        self.world, self.nlocal = w.value, n.value

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        rc = HmmeLib.get().L.hmme_group_unique_id(buf, 128)
        if rc != 0:
            raise HmmeError(rc, HmmeLib.get().L.hmme_group_last_error(None).decode())
        return bytes(buf)

    def _chk(self, rc):
        if rc != 0:
            raise HmmeError(rc, self.lib.L.hmme_group_last_error(self.g).decode())

    def close(self):
        if getattr(self, "g", None):
            self.lib.L.hmme_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_lambda_q16(self, v):
        self._chk(self.lib.L.hmme_group_set_lambda_q16(self.g, C.c_uint32(int(v))))

    def configure(self, width, height, margin_x, margin_y, ref_dist=0):
        self._chk(self.lib.L.hmme_group_configure(self.g, int(width), int(height), int(margin_x), int(margin_y), int(ref_dist)))

    @staticmethod
    def _origin(host, ox, oy):
        assert host.ndim == 2 and host.flags.c_contiguous and host.dtype in (np.int16, np.uint8)
        return C.c_void_p(host.ctypes.data + int(oy * host.shape[1] + ox) * host.itemsize)

    def bind_frame(self, slot, cur, cur_origin, ref, ref_origin, jobs, rng, outs):
        """Zero-argument callable = hmme_group_search_frame_async with every argument marshalled once (launch-bound loops).  cur / ref:
        host planes (int16 or uint8, same type) whose picture sample (0,0) is at [origin_y, origin_x] = *_origin[::-1]."""
        assert cur.dtype == ref.dtype
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        args = (self.g, C.c_int(slot), self._origin(cur, *cur_origin), C.c_int(cur.shape[1]), self._origin(ref, *ref_origin), C.c_int(ref.shape[1]),
                C.c_int(cur.itemsize), C.c_void_p(jobs.ctypes.data), C.c_int(jobs.shape[0]), C.c_int(int(rng)), *[C.c_void_p(o.ctypes.data) for o in outs])
        fn, chk = self.lib.L.hmme_group_search_frame_async, self._chk

        def call():
            chk(fn(*args))
        call.keepalive = (cur, ref, jobs, outs, args)
        return call

    def search_frame_async(self, slot, cur, cur_origin, ref, ref_origin, jobs, rng, outs):
        f = self.bind_frame(slot, cur, cur_origin, ref, ref_origin, jobs, rng, outs)
        f()
        self._keep = f

    def sync(self, slot=-1):
        self._chk(self.lib.L.hmme_group_sync(self.g, int(slot)))

    def bind_sync(self, slot):
        fn, chk, g, s = self.lib.L.hmme_group_sync, self._chk, self.g, C.c_int(slot)
        return lambda: chk(fn(g, s))

    def search_frame(self, cur, cur_origin, ref, ref_origin, jobs, rng):
        jobs = np.ascontiguousarray(jobs, np.int32).reshape(-1, 4)
        outs = MotionEstimator._outs(jobs.shape[0])
        self.search_frame_async(0, cur, cur_origin, ref, ref_origin, jobs, rng, outs)
        self.sync(0)
        return tuple(outs)

    def pipeline_depth(self, njobs, rng):
        """Frames to keep in flight (slots to cycle through) for frames of njobs jobs at +-rng: 2 or 3 (hmme_group_pipeline_depth)."""
        return int(self.lib.L.hmme_group_pipeline_depth(self.g, int(njobs), int(rng)))

    def band(self, njobs, local_index=0):
        f, n = C.c_int(), C.c_int()
        self._chk(self.lib.L.hmme_group_band(self.g, int(njobs), int(local_index), C.byref(f), C.byref(n)))
        return f.value, n.value

    def context(self, local_index=0, slot=0):
        """The per-GPU MotionEstimator behind a slot (borrowed: do not close)."""
        h = self.lib.L.hmme_group_context(self.g, int(local_index), int(slot))
        if not h:
            raise HmmeError(-1, "no such context")
        return MotionEstimator.borrowed(h, self)

    def last_kernel_ms(self, slot=0):
        ms = C.c_float()
        self._chk(self.lib.L.hmme_group_last_kernel_ms(self.g, int(slot), C.byref(ms)))
        return ms.value

    @property
    def kernel_launches(self):
        return int(self.lib.L.hmme_group_kernel_launches(self.g))


class TEncOpenCL:
    """Python mirror of the reference class (TEncOpenCL.h:105-123) with the reference's call order
    findDevice -> compileKernelSource -> createBuffers -> setEnabled(true) (TEncTop.cpp:1129-1145).
    Init methods return bool like the reference; there is no silent CPU fallback -- a failed init
    leaves the object disabled and calcMotionVectors raises."""

    def __init__(self):
        self.deviceId = 0
        self.enabled = False            # the reference leaves this uninitialised
        self._me = None
        self._device_ok = False
        self._kernel_ok = False
        self._lambda = 0.0
        self._info = ""
        self._x = np.zeros(NUM_CTU_PARTS, np.int32)
        self._y = np.zeros(NUM_CTU_PARTS, np.int32)
        self._rui = np.zeros(NUM_CTU_PARTS, np.uint32)
        self._min = np.full(NUM_CTU_PARTS, 0xFFFFFFFF, np.uint32)
        self.last_error = ""

    def findDevice(self, device):
        try:
            n = HmmeLib.get().device_count()
        except HmmeError as e:
            self.last_error = str(e)
            print("ERROR: %s" % e)
            return False
        if device < 0 or device > n - 1:    # TEncOpenCL.cpp:111-115
            device = 0
            print("ID device not found, use default GPU device ")
        self.deviceId = device
        self._device_ok = n > 0
        return self._device_ok

    def compileKernelSource(self, fileName, kernelNameCalc):
        """Kernels are precompiled sm_100a code: the file name is accepted (it must be non-NULL, TEncTop.cpp:1131)
        and otherwise ignored; only the 593-partition kernel exists (AMP_ENC_SPEEDUP=0)."""
        if fileName is None or kernelNameCalc != "calcSAD_AMP":
            self.last_error = "only calcSAD_AMP (593 partitions) is implemented"
            return False
        self._kernel_ok = True
        return True

    def createBuffers(self, maxCtuWidth, maxCtuHeight, searchRange):
        if not (self._device_ok and self._kernel_ok):
            return False
        if maxCtuWidth != 64 or maxCtuHeight != 64:
            self.last_error = "only 64x64 CTUs"
            return False
        try:
            self._me = MotionEstimator(self.deviceId, searchRange)
        except HmmeError as e:
            self.last_error = str(e)
            print("ERROR: %s" % e)
            return False
        self._info = self._me.device_name
        self._me.set_lambda(self._lambda)
        print("Using GPU device              : %s" % self._info)
        return True

    def calcMotionVectors(self, pelCtu, refPlane, ctuPosInPlane, iAreaSize, mvSrchRngLT):
        """pelCtu: (64,64) int16; refPlane: padded int16 plane; ctuPosInPlane = (x, y) array index of the CTU's
        top-left sample (the reference passes that pointer as pelSearch); mvSrchRngLT = (hor, ver)."""
        if not (self.enabled and self._me):
            raise HmmeError(-101, "TEncOpenCL is not initialised/enabled (no CPU fallback)")
        x, y = ctuPosInPlane
        self._x, self._y, self._rui, self._min = self._me.search_ctu(pelCtu, refPlane, x, y, 0, 0, iAreaSize,
                                                                     mvSrchRngLT[0], mvSrchRngLT[1])

    def getDeviceId(self):
        return self.deviceId

    def setDeviceId(self, i):
        self.deviceId = i

    def getDeviceInfo(self):
        return self._info

    def getRuiCost(self):
        return self._rui

    def getX(self):
        return self._x

    def getY(self):
        return self._y

    def setLambda(self, lam):
        self._lambda = lam
        if self._me:
            self._me.set_lambda(lam)

    def setEnabled(self, e):
        self.enabled = bool(e)
