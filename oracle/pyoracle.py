"""ctypes front ends for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module (see oracle/jda_oracle.c header).

  Oracle  -- our restatement, oracle/libjda_oracle.so (instrumented)
  RefLib  -- the reference's own c/jda.c, oracle/_ref/libjda_ref.so, reached
             through the reference's C API (c/jda.h:18-68) and nothing else
  OracleCpp -- our restatement of the double-precision C++ detector, oracle/libjda_oracle_cpp.so
  RefCpp  -- the reference's own C++ detector (src/jda/cascador.cpp, cart.cpp + the detect-path functions of
             data.cpp / btcart.cpp / common.cpp) built against oracle/cvshim/ into oracle/_ref_cpp/, reached
             through oracle/ref_cpp_driver.cpp
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libjda_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libjda_ref.so")


def build(quiet=True):
    """make -C oracle (restatement always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Result(C.Structure):
    # c/jda.h:18-24
    _fields_ = [("n", C.c_int), ("landmark_n", C.c_int),
                ("bboxes", C.POINTER(C.c_int)),
                ("shapes", C.POINTER(C.c_float)),
                ("scores", C.POINTER(C.c_float))]


def _unpack(res, free):
    n, L = res.n, res.landmark_n
    if n > 0:
        boxes = np.ctypeslib.as_array(res.bboxes, shape=(n, 3)).copy()
        shapes = np.ctypeslib.as_array(res.shapes, shape=(n, 2 * L)).copy()
        scores = np.ctypeslib.as_array(res.scores, shape=(n,)).copy()
    else:
        boxes = np.zeros((0, 3), np.int32)
        shapes = np.zeros((0, 2 * L), np.float32)
        scores = np.zeros((0,), np.float32)
    free(res)
    return boxes, scores, shapes


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a, a.ctypes.data_as(C.POINTER(C.c_ubyte)), a.shape[1], a.shape[0]


class RefLib:
    """The unmodified reference library, via its public C API only."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        for f in ("jdaCascadorCreateDouble", "jdaCascadorCreateFloat"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_char_p]
        L.jdaCascadorSerializeTo.argtypes = [C.c_void_p, C.c_char_p]
        L.jdaCascadorSerializeTo.restype = None
        L.jdaCascadorRelease.argtypes = [C.c_void_p]
        L.jdaCascadorRelease.restype = None
        L.jdaDetect.restype = _Result
        L.jdaDetect.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int,
                                C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]
        L.jdaResultRelease.argtypes = [_Result]
        L.jdaResultRelease.restype = None

    def load(self, path, double=True):
        f = self.lib.jdaCascadorCreateDouble if double else self.lib.jdaCascadorCreateFloat
        return f(os.fsencode(path))

    def save_f32(self, handle, path):
        self.lib.jdaCascadorSerializeTo(handle, os.fsencode(path))

    def release(self, handle):
        self.lib.jdaCascadorRelease(handle)

    def detect(self, handle, img, scale=1.25, step=0.1, min_size=24, max_size=-1, th=0.0):
        a, p, w, h = _img(img)
        res = self.lib.jdaDetect(handle, p, w, h, scale, step, min_size, max_size, th)
        return _unpack(res, self.lib.jdaResultRelease)


class Oracle:
    """Our instrumented restatement."""

    N_STATS = 19

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.jdo_load.restype = C.c_void_p
        L.jdo_load.argtypes = [C.c_char_p, C.c_int]
        L.jdo_free.argtypes = [C.c_void_p]
        L.jdo_free.restype = None
        L.jdo_save_f32.argtypes = [C.c_void_p, C.c_char_p]
        L.jdo_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.jdo_dims.restype = None
        L.jdo_resize.argtypes = [C.POINTER(C.c_ubyte), C.c_int, C.c_int,
                                 C.POINTER(C.c_ubyte), C.c_int, C.c_int]
        L.jdo_resize.restype = None
        L.jdo_levels.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                 C.POINTER(C.c_int), C.c_int]
        L.jdo_nms.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                              C.POINTER(C.c_ubyte)]
        L.jdo_nms.restype = None
        L.jdo_detect.restype = _Result
        L.jdo_detect.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int,
                                 C.c_float, C.c_float, C.c_int, C.c_int, C.c_float]
        L.jdo_detect_raw.restype = _Result
        L.jdo_detect_raw.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                     C.POINTER(C.c_longlong)]
        L.jdo_detect_raw_k.restype = _Result
        L.jdo_detect_raw_k.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int,
                                       C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_longlong)]
        L.jdo_trace_k.restype = C.c_longlong
        L.jdo_trace_k.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_float,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_float), C.POINTER(C.c_ubyte),
                                  C.c_longlong, C.c_longlong]
        L.jdo_result_free.argtypes = [_Result]
        L.jdo_result_free.restype = None
        L.jdo_trace.restype = C.c_longlong
        L.jdo_trace.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_float,
                                C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                C.POINTER(C.c_float), C.POINTER(C.c_ubyte),
                                C.c_longlong, C.c_longlong]
        L.jdo_count_windows.restype = C.c_longlong
        L.jdo_count_windows.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]

    def load(self, path, double=True):
        return self.lib.jdo_load(os.fsencode(path), 1 if double else 0)

    def release(self, h):
        self.lib.jdo_free(h)

    def save_f32(self, h, path):
        return self.lib.jdo_save_f32(h, os.fsencode(path))

    def dims(self, h):
        d = (C.c_int * 4)()
        self.lib.jdo_dims(h, d)
        return tuple(d)

    def resize(self, img, dw, dh):
        a, p, w, h = _img(img)
        out = np.empty((dh, dw), np.uint8)
        self.lib.jdo_resize(p, w, h, out.ctypes.data_as(C.POINTER(C.c_ubyte)), dw, dh)
        return out

    def levels(self, w, h, scale=1.25, min_size=24, max_size=-1):
        buf = (C.c_int * 256)()
        n = self.lib.jdo_levels(w, h, scale, min_size, max_size, buf, 256)
        return list(buf[:min(n, 256)])

    def count_windows(self, w, h, scale=1.25, min_size=24, max_size=-1):
        return int(self.lib.jdo_count_windows(w, h, scale, min_size, max_size))

    def nms(self, boxes, scores):
        boxes = np.ascontiguousarray(boxes, np.int32)
        scores = np.ascontiguousarray(scores, np.float32)
        keep = np.zeros(max(len(scores), 1), np.uint8)
        self.lib.jdo_nms(len(scores), boxes.ctypes.data_as(C.POINTER(C.c_int)),
                         scores.ctypes.data_as(C.POINTER(C.c_float)),
                         keep.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return keep[:len(scores)].astype(bool)

    def detect(self, h, img, scale=1.25, step=0.1, min_size=24, max_size=-1, th=0.0):
        a, p, w, hh = _img(img)
        res = self.lib.jdo_detect(h, p, w, hh, scale, step, min_size, max_size, th)
        return _unpack(res, self.lib.jdo_result_free)

    def detect_raw(self, h, img, scale=1.25, min_size=24, max_size=-1, th=0.0,
                   t_limit=0, use_th=True, k_limit=0):
        """pre-NMS hits (scan order, normalised shapes) + stats dict.
        k_limit > 0: t_limit full stages (0 allowed), then carts [0, k_limit) of stage t_limit, no regression
        after them (Validate's unfinished stage, src/jda/cascador.cpp:199-209)."""
        a, p, w, hh = _img(img)
        st = (C.c_longlong * self.N_STATS)()
        res = self.lib.jdo_detect_raw_k(h, p, w, hh, scale, min_size, max_size, th,
                                        t_limit, k_limit, 1 if use_th else 0, st)
        boxes, scores, shapes = _unpack(res, self.lib.jdo_result_free)
        stats = {"windows": st[0], "carts": st[1], "ub_reads": st[2],
                 "stage_survivors": list(st[3:3 + 16])}
        return boxes, scores, shapes, stats

    def trace(self, h, img, scale=1.25, min_size=24, max_size=-1, t_limit=0, leaf_range=None, k_limit=0):
        """per-window (carts evaluated, exit score) in scan order; optional leaves."""
        a, p, w, hh = _img(img)
        nwin = self.count_windows(w, hh, scale, min_size, max_size)
        tn = np.zeros(max(nwin, 1), np.int32)
        ts = np.zeros(max(nwin, 1), np.float32)
        T, K, _, _ = self.dims(h)
        if leaf_range is not None:
            w0, w1 = leaf_range
            lv = np.full((max(w1 - w0, 1), T * K), 255, np.uint8)
            lp = lv.ctypes.data_as(C.POINTER(C.c_ubyte))
        else:
            w0 = w1 = 0
            lv, lp = None, None
        n = self.lib.jdo_trace_k(h, p, w, hh, scale, min_size, max_size, t_limit, k_limit,
                               tn.ctypes.data_as(C.POINTER(C.c_int)),
                               ts.ctypes.data_as(C.POINTER(C.c_float)), lp, w0, w1)
        assert n == nwin, (n, nwin)
        return tn[:nwin], ts[:nwin], lv


REF_CPP_SO = os.path.join(HERE, "_ref_cpp", "libjda_ref_cpp.so")
REF_CPP_RUN = os.path.join(HERE, "_ref_cpp", "run")      # its parent holds config.json (Config reads "../config.json")


class RefCpp:
    """The reference's own JoinCascador (double precision, fddb.method = 1) behind oracle/ref_cpp_driver.cpp.
    Models must be double-flavour files (JoinCascador::SerializeFrom reads doubles)."""

    def __init__(self, path=REF_CPP_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        os.makedirs(REF_CPP_RUN, exist_ok=True)
        L = self.lib = C.CDLL(path)
        L.jref_open.restype = C.c_void_p
        L.jref_open.argtypes = [C.c_char_p, C.c_char_p]
        L.jref_close.argtypes = [C.c_void_p]
        L.jref_close.restype = None
        L.jref_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.jref_dims.restype = None
        L.jref_detect.restype = C.c_int
        L.jref_detect.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_double, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int)),
                                  C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)),
                                  C.POINTER(C.c_double)]
        L.jref_release.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.jref_release.restype = None
        L.jref_trace.restype = C.c_longlong
        L.jref_trace.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_longlong]

    def load(self, path):
        return self.lib.jref_open(os.fsencode(path), os.fsencode(REF_CPP_RUN))

    def release(self, h):
        self.lib.jref_close(h)

    def dims(self, h):
        d = (C.c_int * 6)()
        self.lib.jref_dims(h, d)
        return dict(zip(("T", "K", "L", "depth", "stage", "cart"), d))

    def set_shift(self, shift_size=0.0, tick=0):
        """face.random_shift with the tick that seeds RandomShape's RNG fixed (0: real ticks, shift_size 0 = what
        src/test.cpp forces); returns the (x, y) the reference's RNG then draws for every window."""
        self.lib.jref_set_shift.argtypes = [C.c_double, C.c_longlong, C.POINTER(C.c_double)]
        self.lib.jref_set_shift.restype = None
        xy = (C.c_double * 2)()
        self.lib.jref_set_shift(float(shift_size), int(tick), xy)
        return float(xy[0]), float(xy[1])

    def detect(self, h, img, minimum_size=20, step=5, scale=1.2, overlap=0.3, nms=True, similarity=False):
        """JoinCascador::Detect: (rects[n,4], scores[n] f64, shapes[n,2L] f64, stats dict)"""
        a, p, w, hh = _img(img)
        r, s, sh = C.POINTER(C.c_int)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        st = (C.c_double * 4)()
        n = self.lib.jref_detect(h, p, w, hh, minimum_size, step, scale, overlap, 1 if nms else 0, 1 if similarity else 0,
                                 C.byref(r), C.byref(s), C.byref(sh), st)
        D = 2 * self.dims(h)["L"]
        if n > 0:
            out = (np.ctypeslib.as_array(r, shape=(n, 4)).copy(), np.ctypeslib.as_array(s, shape=(n,)).copy(),
                   np.ctypeslib.as_array(sh, shape=(n, D)).copy())
        else:
            out = (np.zeros((0, 4), np.int32), np.zeros((0,), np.float64), np.zeros((0, D), np.float64))
        self.lib.jref_release(r, s, sh)
        stats = {"patch_n": int(st[0]), "face_patch_n": int(st[1]), "nonface_patch_n": int(st[2]),
                 "cart_gothrough_n": int(st[3])}
        return out + (stats,)

    def trace(self, h, img, minimum_size=20, step=5, scale=1.2, similarity=False):
        """Validate on every window in detectMultiScale1's order: (carts evaluated, exit score)"""
        a, p, w, hh = _img(img)
        sim = 1 if similarity else 0
        n = self.lib.jref_trace(h, p, w, hh, minimum_size, step, scale, sim, None, None, 0)
        tn = np.zeros(max(n, 1), np.int32)
        ts = np.zeros(max(n, 1), np.float64)
        m = self.lib.jref_trace(h, p, w, hh, minimum_size, step, scale, sim, tn.ctypes.data_as(C.POINTER(C.c_int)),
                                ts.ctypes.data_as(C.POINTER(C.c_double)), n)
        assert m == n
        return tn[:n], ts[:n]


ORACLE_CPP_SO = os.path.join(HERE, "libjda_oracle_cpp.so")


class OracleCpp:
    """Restatement of the reference's double-precision C++ detector (JoinCascador::Detect, fddb.method = 1),
    oracle/jda_oracle_cpp.c.  Pinned bit for bit against RefCpp (tests/test_oracle_cpp.py)."""

    # model/config.json "fddb": minimum_size 20, step 5, scale 1.2, overlap 0.3, nms true
    DEFAULTS = dict(minimum_size=20, step=5, scale=1.2, overlap=0.3, nms=True)

    def __init__(self, path=ORACLE_CPP_SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.jcpp_load.restype = C.c_void_p
        L.jcpp_load.argtypes = [C.c_char_p, C.c_int]
        L.jcpp_free.argtypes = [C.c_void_p]
        L.jcpp_free.restype = None
        L.jcpp_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.jcpp_dims.restype = None
        L.jcpp_levels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.c_int]
        L.jcpp_count_windows.restype = C.c_longlong
        L.jcpp_count_windows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.jcpp_nms.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_int)]
        L.jcpp_detect.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_double, C.c_int, C.POINTER(C.POINTER(C.c_int)),
                                  C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)),
                                  C.POINTER(C.c_longlong)]
        L.jcpp_release.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.jcpp_release.restype = None
        L.jcpp_trace.restype = C.c_longlong
        L.jcpp_trace.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.POINTER(C.c_int), C.POINTER(C.c_double)]

    def load(self, path, double=True):
        return self.lib.jcpp_load(os.fsencode(path), 1 if double else 0)

    def set_options(self, similarity=False, shift=(0.0, 0.0)):
        """face.similarity_transform and the initial shift (x, y) RandomShape adds to the mean shape; process-wide,
        like the reference's Config singleton.  Defaults = what the shipped config / src/test.cpp run with."""
        self.lib.jcpp_set_options.argtypes = [C.c_int, C.c_double, C.c_double]
        self.lib.jcpp_set_options.restype = None
        self.lib.jcpp_set_options(1 if similarity else 0, float(shift[0]), float(shift[1]))

    def release(self, h):
        self.lib.jcpp_free(h)

    def dims(self, h):
        d = (C.c_int * 7)()
        self.lib.jcpp_dims(h, d)
        return dict(zip(("T", "K", "L", "depth", "stage", "cart", "any_scaled"), d))

    def levels(self, w, h, minimum_size=20, scale=1.2):
        buf = (C.c_int * 256)()
        n = self.lib.jcpp_levels(w, h, minimum_size, scale, buf, 256)
        return list(buf[:min(n, 256)])

    def count_windows(self, w, h, minimum_size=20, step=5, scale=1.2):
        return int(self.lib.jcpp_count_windows(w, h, minimum_size, step, scale))

    def nms(self, rects, scores, overlap=0.3):
        rects = np.ascontiguousarray(rects, np.int32)
        scores = np.ascontiguousarray(scores, np.float64)
        picked = np.zeros(max(len(scores), 1), np.int32)
        n = self.lib.jcpp_nms(len(scores), rects.ctypes.data_as(C.POINTER(C.c_int)),
                              scores.ctypes.data_as(C.POINTER(C.c_double)), overlap,
                              picked.ctypes.data_as(C.POINTER(C.c_int)))
        return picked[:max(n, 0)].copy()

    def detect(self, h, img, minimum_size=20, step=5, scale=1.2, overlap=0.3, nms=True):
        """(rects[n,4] i32 = x y w h, scores[n] f64, shapes[n,2L] f64 in image pixels, carts evaluated)"""
        a, p, w, hh = _img(img)
        r, s, sh = C.POINTER(C.c_int)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        carts = C.c_longlong(0)
        n = self.lib.jcpp_detect(h, p, w, hh, minimum_size, step, scale, overlap, 1 if nms else 0,
                                 C.byref(r), C.byref(s), C.byref(sh), C.byref(carts))
        if n < 0:
            raise RuntimeError("jcpp_detect refused the model or the arguments")
        D = 2 * self.dims(h)["L"]
        if n > 0:
            out = (np.ctypeslib.as_array(r, shape=(n, 4)).copy(), np.ctypeslib.as_array(s, shape=(n,)).copy(),
                   np.ctypeslib.as_array(sh, shape=(n, D)).copy())
        else:
            out = (np.zeros((0, 4), np.int32), np.zeros((0,), np.float64), np.zeros((0, D), np.float64))
        self.lib.jcpp_release(r, s, sh)
        return out + (int(carts.value),)

    def trace(self, h, img, minimum_size=20, step=5, scale=1.2):
        a, p, w, hh = _img(img)
        nwin = self.count_windows(w, hh, minimum_size, step, scale)
        tn = np.zeros(max(nwin, 1), np.int32)
        ts = np.zeros(max(nwin, 1), np.float64)
        n = self.lib.jcpp_trace(h, p, w, hh, minimum_size, step, scale, tn.ctypes.data_as(C.POINTER(C.c_int)),
                                ts.ctypes.data_as(C.POINTER(C.c_double)))
        assert n == nwin, (n, nwin)
        return tn[:nwin], ts[:nwin]
