"""Python front end of libjda_b200.so -- mirrors the reference's C API (c/jda.h:18-68).

    c = Cascador("model.bin", double=True)          # jdaCascadorCreateDouble / ...Float
    boxes, scores, shapes = c.detect(gray, scale=1.25, step=0.1, min_size=24, max_size=-1, th=0.0)
    c.save_f32("model_f32.bin")                      # jdaCascadorSerializeTo
    c.close()                                        # jdaCascadorRelease

plus the additive batch / device-resident / trace entry points of include/jda_b200.h.
Every call goes through the C ABI with plain pointers; there is no Python or CPU implementation
of the detect path here -- if the CUDA library is missing or no GPU is visible the call raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("JDA_B200_LIB", "libjda_b200.so"))

DEVICE_INPUT, RAW_HITS, NO_FINAL_TH, NO_TMA, NO_STAGE0_SCAN = 1, 2, 4, 8, 16
SAVE_STAGE_T, SAVE_DOUBLE = 1, 2


class _Result(C.Structure):
    _fields_ = [("n", C.c_int), ("landmark_n", C.c_int),
                ("bboxes", C.POINTER(C.c_int)),
                ("shapes", C.POINTER(C.c_float)),
                ("scores", C.POINTER(C.c_float))]


class Batch(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("width", C.c_int), ("height", C.c_int), ("pitch", C.c_int),
                ("frame_stride", C.c_size_t), ("scale", C.c_float), ("min_size", C.c_int),
                ("max_size", C.c_int), ("th", C.c_float), ("t_limit", C.c_int), ("flags", C.c_int),
                ("k_limit", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("windows", C.c_longlong), ("stage0_survivors", C.c_longlong),
                ("raw_hits", C.c_longlong), ("detections", C.c_longlong),
                ("ms_h2d", C.c_float), ("ms_resize", C.c_float), ("ms_scan", C.c_float),
                ("ms_cascade", C.c_float), ("ms_d2h", C.c_float), ("ms_host", C.c_float),
                ("scan_launches", C.c_int), ("cascade_launches", C.c_int), ("resize_launches", C.c_int),
                ("n_levels", C.c_int), ("levels_smem", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ResultF64(C.Structure):
    _fields_ = [("n", C.c_int), ("landmark_n", C.c_int), ("rects", C.POINTER(C.c_int)),
                ("scores", C.POINTER(C.c_double)), ("shapes", C.POINTER(C.c_double))]


class CppParams(C.Structure):
    """the fddb.* keys JoinCascador::Detect reads (src/jda/common.cpp:178-188); defaults = model/config.json"""
    _fields_ = [("minimum_size", C.c_int), ("step", C.c_int), ("scale", C.c_double), ("overlap", C.c_double),
                ("nms", C.c_int), ("flags", C.c_int), ("similarity_transform", C.c_int),
                ("shift_x", C.c_double), ("shift_y", C.c_double)]


class FlatResult(C.Structure):
    _fields_ = [("n_frames", C.c_int), ("total", C.c_int), ("landmark_n", C.c_int), ("counts", C.POINTER(C.c_int)),
                ("bboxes", C.POINTER(C.c_int)), ("scores", C.POINTER(C.c_float)), ("shapes", C.POINTER(C.c_float))]


class Frame(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("pitch", C.c_int)]


_FRAME_DTYPE = np.dtype([("data", np.uint64), ("width", np.int32), ("height", np.int32), ("pitch", np.int32),
                         ("_pad", np.int32)])

_RESULT_DTYPE = np.dtype([("n", np.int32), ("landmark_n", np.int32), ("bboxes", np.uint64), ("shapes", np.uint64),
                          ("scores", np.uint64)])

_lib = None


def lib():
    """dlopen libjda_b200.so (built in-tree by `make -C jda_b200` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with __graft_entry__.build() or `make -C jda_b200`; "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, cp, ci, cf = C.c_void_p, C.c_char_p, C.c_int, C.c_float
    ub = C.POINTER(C.c_ubyte)
    L.jdaCascadorCreateDouble.restype = vp
    L.jdaCascadorCreateDouble.argtypes = [cp]
    L.jdaCascadorCreateFloat.restype = vp
    L.jdaCascadorCreateFloat.argtypes = [cp]
    L.jdaCascadorSerializeTo.restype = None
    L.jdaCascadorSerializeTo.argtypes = [vp, cp]
    L.jdaCascadorRelease.restype = None
    L.jdaCascadorRelease.argtypes = [vp]
    L.jdaDetect.restype = _Result
    L.jdaDetect.argtypes = [vp, ub, ci, ci, cf, cf, ci, ci, cf]
    L.jdaResultRelease.restype = None
    L.jdaResultRelease.argtypes = [_Result]
    L.jdaB200DetectBatch.restype = ci
    L.jdaB200DetectBatch.argtypes = [vp, vp, C.POINTER(Batch), C.POINTER(_Result), C.POINTER(Stats)]
    L.jdaB200DetectBatchFlat.restype = ci
    L.jdaB200DetectBatchFlat.argtypes = [vp, vp, C.POINTER(Batch), C.POINTER(FlatResult), C.POINTER(Stats)]
    L.jdaB200Submit.restype = ci
    L.jdaB200Submit.argtypes = [vp, vp, C.POINTER(Batch)]
    L.jdaB200Collect.restype = ci
    L.jdaB200Collect.argtypes = [vp, ci, C.POINTER(FlatResult), C.POINTER(Stats)]
    L.jdaB200FlatResultRelease.restype = None
    L.jdaB200FlatResultRelease.argtypes = [C.POINTER(FlatResult)]
    L.jdaB200DetectMixed.restype = ci
    L.jdaB200DetectMixed.argtypes = [vp, C.POINTER(Frame), ci, cf, ci, ci, cf, ci, ci, C.POINTER(_Result),
                                     C.POINTER(Stats)]
    L.jdaB200JoinCascadorDetect.restype = ci
    L.jdaB200JoinCascadorDetect.argtypes = [vp, vp, ci, ci, ci, C.POINTER(CppParams), C.POINTER(ResultF64),
                                            C.POINTER(Stats)]
    L.jdaB200ResultF64Release.restype = None
    L.jdaB200ResultF64Release.argtypes = [C.POINTER(ResultF64), ci]
    L.jdaB200JoinCascadorTrace.restype = C.c_longlong
    L.jdaB200JoinCascadorTrace.argtypes = [vp, ub, ci, ci, C.POINTER(CppParams), C.POINTER(ci), C.POINTER(C.c_double)]
    L.jdaB200JoinCascadorFilterMargins.restype = ci
    L.jdaB200JoinCascadorFilterMargins.argtypes = [vp, C.POINTER(C.c_double), ci]
    L.jdaB200JoinCascadorLevels.restype = ci
    L.jdaB200JoinCascadorLevels.argtypes = [ci, ci, ci, C.c_double, C.POINTER(ci), ci]
    L.jdaB200ResultsRelease.restype = None
    L.jdaB200ResultsRelease.argtypes = [C.POINTER(_Result), ci]
    L.jdaB200SetDevice.restype = ci
    L.jdaB200SetDevice.argtypes = [vp, ci]
    L.jdaB200SetStream.restype = ci
    L.jdaB200SetStream.argtypes = [vp, vp]
    L.jdaB200ModelDims.restype = None
    L.jdaB200ModelDims.argtypes = [vp, C.POINTER(ci)]
    L.jdaB200CoalescingStats.restype = None
    L.jdaB200CoalescingStats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(ci)]
    L.jdaB200LastError.restype = cp
    L.jdaB200LastError.argtypes = []
    L.jdaB200DeviceCount.restype = ci
    L.jdaB200DeviceCount.argtypes = []
    L.jdaB200Levels.restype = ci
    L.jdaB200Levels.argtypes = [ci, ci, cf, ci, ci, C.POINTER(ci), ci]
    L.jdaB200CountWindows.restype = C.c_longlong
    L.jdaB200CountWindows.argtypes = [ci, ci, cf, ci, ci]
    L.jdaB200DescribePlan.restype = ci
    L.jdaB200DescribePlan.argtypes = [ci, ci, cf, ci, ci, C.c_char_p, ci]
    L.jdaB200Nms.restype = None
    L.jdaB200Nms.argtypes = [ci, C.POINTER(ci), C.POINTER(cf), ub]
    L.jdaB200Trace.restype = C.c_longlong
    L.jdaB200Trace.argtypes = [vp, ub, ci, ci, cf, ci, ci, ci, ci, C.POINTER(ci), C.POINTER(cf), ub,
                               C.c_longlong, C.c_longlong]
    L.jdaB200TraceK.restype = C.c_longlong
    L.jdaB200TraceK.argtypes = [vp, ub, ci, ci, cf, ci, ci, ci, ci, ci, C.POINTER(ci), C.POINTER(cf), ub,
                                C.c_longlong, C.c_longlong]
    L.jdaB200SerializeTo.restype = ci
    L.jdaB200SerializeTo.argtypes = [vp, cp, ci]
    L.jdaB200Resize.restype = ci
    L.jdaB200Resize.argtypes = [vp, ub, ci, ci, ub, ci, ci]
    _lib = L
    return L


EXPORTS = ["jdaCascadorCreateDouble", "jdaCascadorCreateFloat", "jdaCascadorSerializeTo",
           "jdaCascadorRelease", "jdaDetect", "jdaResultRelease", "jdaB200DetectBatch",
           "jdaB200SetDevice", "jdaB200SetStream", "jdaB200ModelDims", "jdaB200LastError",
           "jdaB200DeviceCount", "jdaB200Levels", "jdaB200CountWindows", "jdaB200Nms",
           "jdaB200Trace", "jdaB200Resize", "jdaB200DescribePlan", "jdaB200ResultsRelease",
           "jdaB200DetectMixed", "jdaB200JoinCascadorDetect", "jdaB200ResultF64Release",
           "jdaB200JoinCascadorTrace", "jdaB200JoinCascadorLevels", "jdaB200JoinCascadorFilterMargins",
           "jdaB200DetectBatchFlat", "jdaB200FlatResultRelease", "jdaB200TraceK", "jdaB200SerializeTo", "jdaB200Submit", "jdaB200Collect",
           "jdaB200CoalescingStats"]


def last_error():
    return lib().jdaB200LastError().decode()


def device_count():
    return lib().jdaB200DeviceCount()


def levels(w, h, scale=1.25, min_size=24, max_size=-1):
    buf = (C.c_int * 64)()
    n = lib().jdaB200Levels(w, h, scale, min_size, max_size, buf, 64)
    return list(buf[:min(n, 64)])


def count_windows(w, h, scale=1.25, min_size=24, max_size=-1):
    return int(lib().jdaB200CountWindows(w, h, scale, min_size, max_size))


def describe_plan(w, h, scale=1.25, min_size=24, max_size=-1, latency=False):
    """scan-kernel tile plan: list of dicts per level (host only).  latency=True: plan of <= 4-frame calls."""
    buf = C.create_string_buffer(4096)
    lib().jdaB200DescribePlan(w, h, scale, min_size, max_size, buf, -4096 if latency else 4096)
    keys = ["win", "step", "nx", "ny", "tw", "th", "box_w", "box_h", "smem", "windows", "span"]
    return [dict(zip(keys, map(int, ln.split()))) for ln in buf.value.decode().splitlines()]


def levels_cpp(w, h, minimum_size=20, scale=1.2):
    """window sizes of the C++ detector's detectMultiScale1 (cascador.cpp:335,372-373)"""
    buf = (C.c_int * 64)()
    n = lib().jdaB200JoinCascadorLevels(w, h, minimum_size, scale, buf, 64)
    return list(buf[:min(n, 64)])


def count_windows_cpp(w, h, minimum_size=20, step=5, scale=1.2):
    return sum(((w - s) // step + 1) * ((h - s) // step + 1) for s in levels_cpp(w, h, minimum_size, scale)) if step > 0 else 0


def nms(boxes, scores):
    boxes = np.ascontiguousarray(boxes, np.int32)
    scores = np.ascontiguousarray(scores, np.float32)
    keep = np.zeros(max(len(scores), 1), np.uint8)
    lib().jdaB200Nms(len(scores), boxes.ctypes.data_as(C.POINTER(C.c_int)),
                     scores.ctypes.data_as(C.POINTER(C.c_float)),
                     keep.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return keep[:len(scores)].astype(bool)


def _unpack(res, D=None):
    L = lib()
    n, lm = res.n, res.landmark_n
    if n < 0:
        raise RuntimeError("jda_b200 detect failed: " + last_error())
    if n > 0:
        out = (np.ctypeslib.as_array(res.bboxes, shape=(n, 3)).copy(),
               np.ctypeslib.as_array(res.scores, shape=(n,)).copy(),
               np.ctypeslib.as_array(res.shapes, shape=(n, 2 * lm)).copy())
    else:
        out = (np.zeros((0, 3), np.int32), np.zeros((0,), np.float32), np.zeros((0, 2 * lm), np.float32))
    L.jdaResultRelease(res)
    return out


class Cascador:
    """Handle on one loaded model (jdaCascador of c/jda.c:142-151 + its device copy)."""

    def __init__(self, path, double=True, device=None):
        L = lib()
        f = L.jdaCascadorCreateDouble if double else L.jdaCascadorCreateFloat
        self._h = f(os.fsencode(path))
        if not self._h:
            raise RuntimeError("cannot load model %r: %s" % (path, last_error()))
        if device is not None:
            if L.jdaB200SetDevice(self._h, int(device)) != 0:
                raise RuntimeError(last_error())
        d = (C.c_int * 4)()
        L.jdaB200ModelDims(self._h, d)
        self.T, self.K, self.L, self.depth = tuple(d)
        self.last_stats = None

    def close(self):
        if getattr(self, "_h", None):
            lib().jdaCascadorRelease(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def save_f32(self, path):
        lib().jdaCascadorSerializeTo(self._h, os.fsencode(path))

    def save(self, path, flags=0):
        """jdaB200SerializeTo: flags 0 = jdaCascadorSerializeTo's bytes; SAVE_STAGE_T | SAVE_DOUBLE = a file the
        reference's C++ loader accepts (cascador.cpp:126-164)."""
        if lib().jdaB200SerializeTo(self._h, os.fsencode(path), flags) != 0:
            raise RuntimeError("jdaB200SerializeTo failed")

    def set_stream(self, cuda_stream_ptr):
        lib().jdaB200SetStream(self._h, C.c_void_p(cuda_stream_ptr))

    def detect(self, img, scale=1.25, step=0.1, min_size=24, max_size=-1, th=0.0):
        """jdaDetect: (boxes[n,3] i32, scores[n] f32, shapes[n,2L] f32 in image pixels)."""
        a = np.ascontiguousarray(img, np.uint8)
        assert a.ndim == 2
        res = lib().jdaDetect(self._h, a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.shape[1], a.shape[0],
                              scale, step, min_size, max_size, th)
        return _unpack(res)

    def coalescing_stats(self):
        """(jdaDetect calls, device batches they were served in, frames of the largest batch) since creation."""
        a, b, m = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
        lib().jdaB200CoalescingStats(self._h, C.byref(a), C.byref(b), C.byref(m))
        return a.value, b.value, m.value

    def detect_batch(self, frames, scale=1.25, min_size=24, max_size=-1, th=0.0, t_limit=0, flags=0,
                     device_ptr=None, shape=None, pitch=None, frame_stride=None, unpack=True, flat=False, k_limit=0):
        """frames: [n,h,w] u8 numpy array (host), or device_ptr + shape=(n,h,w) for frames resident
        in HBM (any allocator: torch .data_ptr(), cudaMalloc ...).  Returns a list of
        (boxes, scores, shapes) per frame, or just the detection count when unpack=False."""
        L = lib()
        if device_ptr is None:
            a = np.ascontiguousarray(frames, np.uint8)
            assert a.ndim == 3
            n, h, w = a.shape
            ptr = a.ctypes.data
            pitch = w if pitch is None else pitch
            frame_stride = pitch * h if frame_stride is None else frame_stride
        else:
            n, h, w = shape
            ptr = int(device_ptr)
            pitch = w if pitch is None else pitch
            frame_stride = pitch * h if frame_stride is None else frame_stride
            flags |= DEVICE_INPUT
        b = Batch(n, w, h, pitch, frame_stride, scale, min_size, max_size, th, t_limit, flags, k_limit)
        if flat:
            # jdaB200DetectBatchFlat: (counts[n], boxes[total,3], scores[total], shapes[total,2L]), frame order
            fr = FlatResult()
            st = Stats()
            rc = L.jdaB200DetectBatchFlat(self._h, C.c_void_p(ptr), C.byref(b), C.byref(fr), C.byref(st))
            self.last_stats = st.as_dict()
            if rc != 0:
                raise RuntimeError("jdaB200DetectBatchFlat failed: " + last_error())
            tot, D = fr.total, 2 * fr.landmark_n
            counts = np.ctypeslib.as_array(fr.counts, shape=(max(n, 1),))[:n].copy()
            if tot:
                out = (counts, np.ctypeslib.as_array(fr.bboxes, shape=(tot, 3)).copy(),
                       np.ctypeslib.as_array(fr.scores, shape=(tot,)).copy(),
                       np.ctypeslib.as_array(fr.shapes, shape=(tot, D)).copy())
            else:
                out = (counts, np.zeros((0, 3), np.int32), np.zeros((0,), np.float32), np.zeros((0, D), np.float32))
            L.jdaB200FlatResultRelease(C.byref(fr))
            return out
        res = (_Result * n)()
        st = Stats()
        rc = L.jdaB200DetectBatch(self._h, C.c_void_p(ptr), C.byref(b), res, C.byref(st))
        self.last_stats = st.as_dict()
        if rc != 0:
            raise RuntimeError("jdaB200DetectBatch failed: " + last_error())
        return self._unpack_results(res, n, unpack)

    def submit(self, frames, scale=1.25, min_size=24, max_size=-1, th=0.0, t_limit=0, flags=0,
               device_ptr=None, shape=None, pitch=None, frame_stride=None, k_limit=0):
        """jdaB200Submit: copy the batch in and launch its kernels; returns a ticket for collect().  `frames` (a
        C-contiguous [n,h,w] u8 array, or device_ptr + shape) must stay alive and unchanged until then."""
        if device_ptr is None:
            a = frames
            assert a.dtype == np.uint8 and a.ndim == 3 and a.flags["C_CONTIGUOUS"]
            n, h, w = a.shape
            ptr = a.ctypes.data
            pitch = w if pitch is None else pitch
            frame_stride = pitch * h if frame_stride is None else frame_stride
        else:
            n, h, w = shape
            ptr = int(device_ptr)
            pitch = w if pitch is None else pitch
            frame_stride = pitch * h if frame_stride is None else frame_stride
            flags |= DEVICE_INPUT
        b = Batch(n, w, h, pitch, frame_stride, scale, min_size, max_size, th, t_limit, flags, k_limit)
        t = lib().jdaB200Submit(self._h, C.c_void_p(ptr), C.byref(b))
        if t < 0:
            raise RuntimeError("jdaB200Submit failed: " + last_error())
        self._pending = getattr(self, "_pending", {})
        self._pending[t] = (frames, n)
        return t

    def collect(self, ticket):
        """jdaB200Collect: (counts[n], boxes[total,3], scores[total], shapes[total,2L]) like detect_batch(flat=True)"""
        L = lib()
        fr = FlatResult()
        st = Stats()
        rc = L.jdaB200Collect(self._h, ticket, C.byref(fr), C.byref(st))
        _, n = self._pending.pop(ticket, (None, 0))
        self.last_stats = st.as_dict()
        if rc != 0:
            raise RuntimeError("jdaB200Collect failed: " + last_error())
        tot, D = fr.total, 2 * fr.landmark_n
        counts = np.ctypeslib.as_array(fr.counts, shape=(max(n, 1),))[:n].copy()
        if tot:
            out = (counts, np.ctypeslib.as_array(fr.bboxes, shape=(tot, 3)).copy(),
                   np.ctypeslib.as_array(fr.scores, shape=(tot,)).copy(),
                   np.ctypeslib.as_array(fr.shapes, shape=(tot, D)).copy())
        else:
            out = (counts, np.zeros((0, 3), np.int32), np.zeros((0,), np.float32), np.zeros((0, D), np.float32))
        L.jdaB200FlatResultRelease(C.byref(fr))
        return out

    def _unpack_results(self, res, n, unpack=True):
        L = lib()
        lm = self.L
        # vectorised view of the jdaResult array (c/jda.h:18-24: two ints + three pointers)
        view = np.frombuffer(res, dtype=_RESULT_DTYPE, count=n)
        counts = view["n"]
        if not unpack:
            tot = int(counts.sum())
            L.jdaB200ResultsRelease(res, n)
            return tot
        empty = (np.zeros((0, 3), np.int32), np.zeros((0,), np.float32), np.zeros((0, 2 * lm), np.float32))
        out = [empty] * n
        for i in np.nonzero(counts > 0)[0]:
            k = int(counts[i])
            out[i] = (np.ctypeslib.as_array(res[i].bboxes, shape=(k, 3)).copy(),
                      np.ctypeslib.as_array(res[i].scores, shape=(k,)).copy(),
                      np.ctypeslib.as_array(res[i].shapes, shape=(k, 2 * lm)).copy())
        L.jdaB200ResultsRelease(res, n)
        return out

    def detect_mixed(self, frames, scale=1.25, min_size=24, max_size=-1, th=0.0, t_limit=0, flags=0, unpack=True):
        """jdaB200DetectMixed: host frames of different sizes (list of 2-D u8 arrays, row stride = any), one
        launch per kernel over a common canvas.  Result i is what jdaDetect returns for frames[i] alone."""
        n = len(frames)
        keep = [f if (f.dtype == np.uint8 and f.ndim == 2 and f.strides[1] == 1 and f.strides[0] >= f.shape[1])
                else np.ascontiguousarray(f, np.uint8) for f in frames]
        # the jdaB200Frame array (pointer, width, height, pitch; 24 bytes each), filled column by column
        tab = np.zeros(max(n, 1), _FRAME_DTYPE)
        if n:
            tab["data"] = [f.__array_interface__["data"][0] for f in keep]
            tab["height"] = [f.shape[0] for f in keep]
            tab["width"] = [f.shape[1] for f in keep]
            tab["pitch"] = [f.strides[0] if f.shape[0] > 1 else f.shape[1] for f in keep]
        arr = tab.ctypes.data_as(C.POINTER(Frame))
        res = (_Result * max(n, 1))()
        st = Stats()
        rc = lib().jdaB200DetectMixed(self._h, arr, n, scale, min_size, max_size, th, t_limit, flags, res, C.byref(st))
        self.last_stats = st.as_dict()
        if rc != 0:
            raise RuntimeError("jdaB200DetectMixed failed: " + last_error())
        return self._unpack_results(res, n, unpack)

    def detect_cpp(self, frames, minimum_size=20, step=5, scale=1.2, overlap=0.3, nms=True, flags=0, similarity=False,
                   shift=(0.0, 0.0)):
        """jdaB200JoinCascadorDetect: the reference's double-precision C++ detector (JoinCascador::Detect,
        fddb.method = 1).  frames: one [h,w] u8 image or a batch [n,h,w].  Returns (rects[k,4] i32 = x y w h,
        scores[k] f64, shapes[k,2L] f64 in image pixels) -- a list of those for a batch."""
        a = np.ascontiguousarray(frames, np.uint8)
        single = a.ndim == 2
        if single:
            a = a[None]
        n, h, w = a.shape
        prm = CppParams(minimum_size, step, scale, overlap, 1 if nms else 0, flags, 1 if similarity else 0,
                        float(shift[0]), float(shift[1]))
        res = (ResultF64 * max(n, 1))()
        st = Stats()
        rc = lib().jdaB200JoinCascadorDetect(self._h, C.c_void_p(a.ctypes.data), n, w, h, C.byref(prm), res, C.byref(st))
        self.last_stats = st.as_dict()
        if rc != 0:
            raise RuntimeError("jdaB200JoinCascadorDetect failed: " + last_error())
        D = 2 * self.L
        out = []
        for i in range(n):
            k = res[i].n
            if k > 0:
                out.append((np.ctypeslib.as_array(res[i].rects, shape=(k, 4)).copy(),
                            np.ctypeslib.as_array(res[i].scores, shape=(k,)).copy(),
                            np.ctypeslib.as_array(res[i].shapes, shape=(k, D)).copy()))
            else:
                out.append((np.zeros((0, 4), np.int32), np.zeros((0,), np.float64), np.zeros((0, D), np.float64)))
        lib().jdaB200ResultF64Release(res, n)
        return out[0] if single else out

    def filter_margins_cpp(self):
        """per-cart margins of the float32 stage-0 prefilter of the double-precision detector (None: no prefilter)"""
        buf = np.zeros(self.K, np.float64)
        n = lib().jdaB200JoinCascadorFilterMargins(self._h, buf.ctypes.data_as(C.POINTER(C.c_double)), self.K)
        if n < 0:
            raise RuntimeError("jdaB200JoinCascadorFilterMargins failed: " + last_error())
        return buf if n > 0 else None

    def detect_cpp_many(self, frames, **kw):
        """detect_cpp on frames of mixed sizes: one batch call per distinct shape, results in input order"""
        groups = {}
        for i, f in enumerate(frames):
            groups.setdefault(tuple(f.shape), []).append(i)
        out = [None] * len(frames)
        for shape, idx in groups.items():
            res = self.detect_cpp(np.stack([np.ascontiguousarray(frames[i], np.uint8) for i in idx]), **kw)
            for i, r in zip(idx, res):
                out[i] = r
        return out

    def trace_cpp(self, img, minimum_size=20, step=5, scale=1.2, similarity=False, shift=(0.0, 0.0)):
        """JoinCascador::Validate per window in scan order: (carts evaluated, exit score f64)"""
        a = np.ascontiguousarray(img, np.uint8)
        h, w = a.shape
        nwin = count_windows_cpp(w, h, minimum_size, step, scale)
        tn = np.zeros(max(nwin, 1), np.int32)
        ts = np.zeros(max(nwin, 1), np.float64)
        prm = CppParams(minimum_size, step, scale, 0.3, 1, 0, 1 if similarity else 0, float(shift[0]), float(shift[1]))
        n = lib().jdaB200JoinCascadorTrace(self._h, a.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, C.byref(prm),
                                           tn.ctypes.data_as(C.POINTER(C.c_int)), ts.ctypes.data_as(C.POINTER(C.c_double)))
        if n < 0:
            raise RuntimeError("jdaB200JoinCascadorTrace failed: " + last_error())
        assert n == nwin, (n, nwin)
        return tn[:nwin], ts[:nwin]

    def detect_many(self, frames, group=False, **kw):
        """frames of mixed sizes (e.g. FDDB-shaped), results in input order.
        Default: one jdaB200DetectMixed call (all frames share each kernel launch).  group=True: the older
        scheme -- frames grouped by shape, one jdaB200DetectBatch call per group."""
        if not group:
            return self.detect_mixed(frames, **kw)
        groups = {}
        for i, f in enumerate(frames):
            groups.setdefault(tuple(f.shape), []).append(i)
        out = [None] * len(frames)
        stats = None
        for shape, idx in groups.items():
            res = self.detect_batch(np.stack([np.ascontiguousarray(frames[i], np.uint8) for i in idx]), **kw)
            for i, r in zip(idx, res):
                out[i] = r
            st = self.last_stats
            if stats is None:
                stats = dict(st)
            else:
                for k in ("windows", "stage0_survivors", "raw_hits", "detections", "ms_h2d", "ms_resize", "ms_scan",
                          "ms_cascade", "ms_d2h", "ms_host", "scan_launches", "cascade_launches", "resize_launches"):
                    stats[k] += st[k]
        self.last_stats = stats
        return out

    def trace(self, img, scale=1.25, min_size=24, max_size=-1, t_limit=0, flags=0, leaf_range=None, k_limit=0):
        """per-window (carts evaluated, exit score) in scan order + optional leaf indices."""
        a = np.ascontiguousarray(img, np.uint8)
        h, w = a.shape
        nwin = count_windows(w, h, scale, min_size, max_size)
        tn = np.zeros(max(nwin, 1), np.int32)
        ts = np.zeros(max(nwin, 1), np.float32)
        if leaf_range is not None:
            w0, w1 = leaf_range
            lv = np.full((max(w1 - w0, 1), self.T * self.K), 255, np.uint8)
            lp = lv.ctypes.data_as(C.POINTER(C.c_ubyte))
        else:
            w0 = w1 = 0
            lv, lp = None, None
        n = lib().jdaB200TraceK(self._h, a.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, scale, min_size,
                               max_size, t_limit, k_limit, flags, tn.ctypes.data_as(C.POINTER(C.c_int)),
                               ts.ctypes.data_as(C.POINTER(C.c_float)), lp, w0, w1)
        if n < 0:
            raise RuntimeError("jdaB200Trace failed: " + last_error())
        assert n == nwin, (n, nwin)
        return tn[:nwin], ts[:nwin], lv

    def resize(self, img, dw, dh):
        a = np.ascontiguousarray(img, np.uint8)
        out = np.empty((dh, dw), np.uint8)
        rc = lib().jdaB200Resize(self._h, a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.shape[1], a.shape[0],
                                 out.ctypes.data_as(C.POINTER(C.c_ubyte)), dw, dh)
        if rc != 0:
            raise RuntimeError("jdaB200Resize failed: " + last_error())
        return out
