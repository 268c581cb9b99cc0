"""The restatement of the reference's double-precision C++ detector (oracle/jda_oracle_cpp.c: JoinCascador::Detect,
fddb.method = 1) and the host logic of its CUDA counterpart.  No GPU needed.

PINNED (round 2): the reference's own src/jda/cascador.cpp and cart.cpp (whole) + the detect-path functions of
data.cpp / btcart.cpp / common.cpp are compiled from where they lie against a ~200-line stand-in for the OpenCV core
headers (oracle/cvshim/, oracle/Makefile target ref_cpp) and the restatement is compared with that binary bit for bit:
JoinCascador::Detect (faces, scores, landmarks, patch and cart statistics), the raw list without NMS, and Validate on
every window (carts evaluated, exit score) -- shipped model, full-precision synthetic models, training snapshots whose
header stops inside a stage.  Also kept: the independent restatements of its pieces (multimap NMS, window ladder, a
from-scratch Python Validate) and the error bound that lets the float32 scan kernel prefilter stage 0 of the double path.
Scope of the pin: fddb.method = 1, models without scale != 0 nodes; face.similarity_transform on and off, zero and
seed-fixed non-zero initial shifts (the last test of this file) -- what runs through cv::resize is OpenCV's arithmetic,
which the stand-in does not reproduce, and cv::norm inside STParameter::Calc is the stand-in's index-order sum.
"""
import numpy as np
import pytest

from jda_b200 import api, synth
from tests.conftest import SHIPPED_F32


@pytest.fixture(scope="module")
def ocpp():
    from oracle import pyoracle
    return pyoracle.OracleCpp()


@pytest.fixture(scope="module")
def ocpp_shipped(ocpp):
    h = ocpp.load(SHIPPED_F32, double=False)
    assert h
    yield h
    ocpp.release(h)


@pytest.fixture(scope="module")
def refcpp():
    """the reference's own C++ detector (oracle/_ref_cpp); skip when it was never built"""
    import os
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_CPP_SO):
        pytest.skip("oracle/_ref_cpp/libjda_ref_cpp.so not built (needs /root/reference)")
    return pyoracle.RefCpp()


@pytest.fixture(scope="module")
def wide_shipped(tmp_path_factory):
    """the shipped model as a double-flavour file (JoinCascador::SerializeFrom reads doubles; f32 -> f64 is exact)"""
    return synth.widen_f32_model(SHIPPED_F32, str(tmp_path_factory.mktemp("wide") / "wide.model"))


def _same64(a, b):
    np.testing.assert_array_equal(np.ascontiguousarray(a, np.float64).view(np.uint64),
                                  np.ascontiguousarray(b, np.float64).view(np.uint64))


PIN_FRAMES = {"faces": lambda: synth.face_canvas(), "facemix": lambda: synth.facemix_frame(3, 320, 240),
              "noise": lambda: synth.noise_frame(5, 160, 120), "fddb_shape": lambda: synth.facemix_frame(7, *synth.fddb_shape(7))}


@pytest.mark.parametrize("frame", list(PIN_FRAMES))
def test_restatement_pinned_to_reference_detect(ocpp, ocpp_shipped, refcpp, wide_shipped, frame):
    """JoinCascador::Detect of the reference binary vs the restatement: faces, scores and landmarks as raw 64-bit
    patterns, with and without NMS, for config.json's fddb settings and two others; patch / cart statistics"""
    hr = refcpp.load(wide_shipped)
    assert hr and refcpp.dims(hr) == dict(T=5, K=540, L=27, depth=4, stage=5, cart=-1)
    img = PIN_FRAMES[frame]()
    for kw in (dict(), dict(minimum_size=30, step=7, scale=1.3, overlap=0.5), dict(minimum_size=24, step=3, scale=1.25, nms=False)):
        if kw.get("step") == 3 and img.shape[1] > 400:
            continue                                   # (keeps the CPU suite short)
        rr, rs, rsh, st = refcpp.detect(hr, img, **kw)
        orr, os_, osh, carts = ocpp.detect(ocpp_shipped, img, **kw)
        np.testing.assert_array_equal(rr, orr)
        _same64(rs, os_)
        _same64(rsh, osh)
        # DetectionStatisic (cascador.hpp:14-25): patches, and carts walked by the windows that were rejected
        raw = ocpp.detect(ocpp_shipped, img, **dict(kw, nms=False))
        mn, stp, sc = kw.get("minimum_size", 20), kw.get("step", 5), kw.get("scale", 1.2)
        assert st["patch_n"] == ocpp.count_windows(img.shape[1], img.shape[0], mn, stp, sc)
        assert st["face_patch_n"] == len(raw[1])
        assert st["cart_gothrough_n"] == carts - 2700 * len(raw[1])
    if frame == "faces":
        assert rr.shape[0] >= 1
    refcpp.release(hr)


def test_restatement_pinned_to_reference_validate_trace(ocpp, ocpp_shipped, refcpp, wide_shipped):
    """Validate per window (cascador.cpp:166-211) through the reference binary: carts evaluated and exit score"""
    hr = refcpp.load(wide_shipped)
    for img in (synth.face_canvas()[20:260, 40:300].copy(), synth.noise_frame(2, 90, 70)):
        rn, rs = refcpp.trace(hr, img)
        on, os_ = ocpp.trace(ocpp_shipped, img)
        np.testing.assert_array_equal(rn, on)
        _same64(rs, os_)
        assert len(rn) > 500
    refcpp.release(hr)


def test_restatement_pinned_to_reference_synthetic_models(ocpp, refcpp, tmp_path):
    """full-precision double models (values that do not survive a float round trip): pass-all, rejecting with many
    normalised carts, and training snapshots whose header stops inside a stage (cascador.cpp:199-209)"""
    img = synth.facemix_frame(11, 120, 100)
    for name, kw, hdr in (("pass", dict(seed=1, mode="passall"), None), ("rej", dict(seed=2, mode="reject", norm_every=7), None),
                          ("snap", dict(seed=4, mode="reject"), (2, 17)), ("snap0", dict(seed=5, mode="reject"), (0, 300)),
                          ("small", dict(seed=6, mode="reject", T=3, K=64, L=9, norm_every=5), None)):
        p = synth.write_model(str(tmp_path / (name + ".model")), **kw)
        if hdr:
            b = bytearray(open(p, "rb").read())
            b[20:24] = hdr[0].to_bytes(4, "little"); b[24:28] = hdr[1].to_bytes(4, "little", signed=True)
            open(p, "wb").write(bytes(b))
        hr, ho = refcpp.load(p), ocpp.load(p, True)
        assert hr and ho
        if hdr:
            assert (refcpp.dims(hr)["stage"], refcpp.dims(hr)["cart"]) == hdr
        rn, rs = refcpp.trace(hr, img)
        on, os_ = ocpp.trace(ho, img)
        np.testing.assert_array_equal(rn, on)
        _same64(rs, os_)
        for nms in (True, False):
            rr, rsc, rsh, _ = refcpp.detect(hr, img, nms=nms)
            orr, osc, osh, _ = ocpp.detect(ho, img, nms=nms)
            np.testing.assert_array_equal(rr, orr)
            _same64(rsc, osc)
            _same64(rsh, osh)
        if name == "pass":
            assert len(rsc) == ocpp.count_windows(120, 100)
        refcpp.release(hr); ocpp.release(ho)


def test_header_and_scope(ocpp, ocpp_shipped, tmp_path):
    d = ocpp.dims(ocpp_shipped)
    assert (d["T"], d["K"], d["L"], d["depth"]) == (5, 540, 27, 4)
    assert (d["stage"], d["cart"]) == (5, -1)      # the float writer's T+1 quirk (c/jda.c:662) normalised
    assert d["any_scaled"] == 0                    # the shipped model never samples the h / q planes
    # a double-flavour file with the same values loads to the same model -> same answers
    wide = synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model"))
    h2 = ocpp.load(wide, double=True)
    img = synth.facemix_frame(3, 200, 150)
    a, b = ocpp.detect(ocpp_shipped, img, nms=False), ocpp.detect(h2, img, nms=False)
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)
    ocpp.release(h2)
    # models with scale != 0 nodes need cv::resize'd planes: refused
    p = synth.write_model(str(tmp_path / "scaled.model"), seed=3, scales=(0, 1, 2), coord_max=0.45)
    h3 = ocpp.load(p, True)
    with pytest.raises(RuntimeError):
        ocpp.detect(h3, img)
    ocpp.release(h3)


@pytest.mark.parametrize("w,h,mn,sc", [(640, 480, 20, 1.2), (450, 333, 20, 1.2), (1920, 1080, 20, 1.2),
                                       (640, 480, 30, 1.3), (19, 100, 20, 1.2), (20, 20, 20, 1.2),
                                       (640, 480, 20, 1.0), (640, 480, 20, 1.04), (640, 480, 0, 1.2)])
def test_window_ladder(ocpp, w, h, mn, sc):
    """detectMultiScale1: win = int(win * factor) from fddb.minimum_size while it fits (cascador.cpp:335,372-373)"""
    want = []
    win = mn
    if mn > 0 and sc > 1.0:
        while win <= w and win <= h:
            want.append(win)
            nxt = int(win * sc)
            if nxt <= win:
                break
            win = nxt
    assert ocpp.levels(w, h, mn, sc) == want == api.levels_cpp(w, h, mn, sc)
    assert ocpp.count_windows(w, h, mn, 5, sc) == api.count_windows_cpp(w, h, mn, 5, sc)


def _multimap_nms(rects, scores, overlap):
    """cascador.cpp:387-429 with a literal multimap stand-in: list of (score, idx) kept sorted, equal keys in
    insertion order; take the last, erase everything with IoU > overlap (itself included)."""
    m = []
    for i, s in enumerate(scores):
        j = len(m)
        while j > 0 and m[j - 1][0] > s:
            j -= 1
        m.insert(j, (s, i))
    picked = []
    while m:
        last = m[-1][1]
        picked.append(last)
        keep = []
        for s, idx in m:
            x1 = max(rects[idx][0], rects[last][0]); y1 = max(rects[idx][1], rects[last][1])
            x2 = min(rects[idx][0] + rects[idx][2], rects[last][0] + rects[last][2])
            y2 = min(rects[idx][1] + rects[idx][3], rects[last][1] + rects[last][3])
            w, h = max(0.0, float(x2 - x1)), max(0.0, float(y2 - y1))
            a1, a2 = float(rects[idx][2] * rects[idx][3]), float(rects[last][2] * rects[last][3])
            if not (w * h / (a1 + a2 - w * h) > overlap):
                keep.append((s, idx))
        m = keep
    return picked


def test_nms_against_literal_multimap(ocpp):
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 9, 60, 250):
        size = rng.integers(20, 120, n)
        rects = np.stack([rng.integers(0, 200, n), rng.integers(0, 200, n), size, size], 1).astype(np.int32)
        scores = rng.normal(0, 1, n)
        if n:
            scores[rng.integers(0, n, n // 3)] = 0.25     # ties: the multimap keeps insertion order
        for ov in (0.3, 0.0, 0.9):
            assert list(ocpp.nms(rects, scores, ov)) == _multimap_nms(rects.tolist(), scores.tolist(), ov)


def test_known_answers_regression_pin(ocpp, ocpp_shipped):
    """what the restatement finds on the reference's own face image with model/config.json's fddb settings
    (minimum_size 20, step 5, scale 1.2, overlap 0.3): pins the restatement against accidental change"""
    img = synth.face_canvas()
    rects, scores, shapes, carts = ocpp.detect(ocpp_shipped, img)
    assert rects.tolist() == [[400, 290, 112, 112], [55, 40, 230, 230]]
    np.testing.assert_allclose(scores, [2.08002624, 2.04335326], rtol=0, atol=5e-9)
    assert shapes.shape == (2, 54) and carts == 4475389
    # landmarks lie inside (a slightly grown copy of) their boxes
    for r, s in zip(rects, shapes):
        assert (s[0::2] > r[0] - 0.2 * r[2]).all() and (s[0::2] < r[0] + 1.2 * r[2]).all()
        assert (s[1::2] > r[1] - 0.2 * r[3]).all() and (s[1::2] < r[1] + 1.2 * r[3]).all()
    raw = ocpp.detect(ocpp_shipped, img, nms=False)
    assert len(raw[1]) == 365 and (np.diff(raw[0][:, 2]) >= 0).all()      # scan order: window size ascending
    tn, ts = ocpp.trace(ocpp_shipped, img)
    assert len(tn) == 140215 == ocpp.count_windows(640, 480) and int(tn.sum()) == carts
    assert int((tn == 2700).sum()) == 365                                    # faces ran every cart of every stage


def _stage0_tables(path=SHIPPED_F32):
    raw = open(path, "rb").read()
    cart = np.frombuffer(raw, np.uint8, 540 * 268, 28 + 216).reshape(540, 268)
    leaf = cart[:, 224:256].copy().view(np.float32).astype(np.float64)
    tms = cart[:, 256:268].copy().view(np.float32).astype(np.float64)
    return leaf, tms


def test_prefilter_margins_bound_float_vs_double_scores():
    """the float32 scan kernel may drop a window at cart k only if its float score is below th_k - margin_k;
    margin_k must bound |float32 score - double score| for every leaf sequence.  Checked on random, extreme and
    alternating leaf paths of the shipped stage 0 (float path = k2_scan's arithmetic, double path = Validate's)."""
    c = api.Cascador(SHIPPED_F32, double=False)
    margins = c.filter_margins_cpp()
    c.close()
    assert margins is not None and len(margins) == 540 and (margins > 0).all() and margins.max() < 0.05
    leaf, tms = _stage0_tables()
    rng = np.random.default_rng(0)
    n = 4000
    paths = rng.integers(0, 8, (n, 540))
    paths[0] = np.abs(leaf).argmax(1); paths[1] = leaf.argmax(1); paths[2] = leaf.argmin(1)
    paths[3] = np.where(np.arange(540) % 2, leaf.argmax(1), leaf.argmin(1))
    s32 = np.zeros(n, np.float32)
    s64 = np.zeros(n, np.float64)
    worst = 0.0
    for k in range(540):
        lf = leaf[k][paths[:, k]]
        s32 = (s32 + lf.astype(np.float32)).astype(np.float32)
        s64 = s64 + lf
        mean, sd = tms[k, 1], tms[k, 2]
        if np.float32(mean) != 0 or np.float32(sd) != 1:
            s32 = ((s32 - np.float32(mean)).astype(np.float32) / np.float32(sd)).astype(np.float32)
        s64 = (s64 - mean) / sd
        err = np.abs(s32.astype(np.float64) - s64).max()
        assert err <= margins[k], (k, err, margins[k])
        worst = max(worst, err / margins[k])
    assert worst < 0.5     # the bound is conservative by construction (factor 4)


def test_synthetic_double_models(ocpp, tmp_path):
    """full-precision double models: pass-all (every window is a face), rejecting, and a training snapshot whose
    header stops inside stage 2 (cascador.cpp:199-209)"""
    img = synth.blur_frame(9, 96, 80)
    p = synth.write_model(str(tmp_path / "pass.model"), seed=1, mode="passall")
    h = ocpp.load(p, True)
    r, s, sh, carts = ocpp.detect(h, img, nms=False)
    nwin = ocpp.count_windows(96, 80)
    assert len(s) == nwin and carts == nwin * 2700
    ocpp.release(h)
    q = synth.write_model(str(tmp_path / "rej.model"), seed=2, mode="reject")
    b = bytearray(open(q, "rb").read())
    b[20:24] = (2).to_bytes(4, "little"); b[24:28] = (17).to_bytes(4, "little", signed=True)
    snap = tmp_path / "snap.model"
    snap.write_bytes(bytes(b))
    h = ocpp.load(str(snap), True)
    d = ocpp.dims(h)
    assert (d["stage"], d["cart"]) == (2, 17)
    tn, ts = ocpp.trace(h, img)
    assert tn.max() <= 2 * 540 + 18
    ocpp.release(h)


def _py_validate(img, x, y, win, mean, nodes, leaf, cart3, w, T, K):
    """second, independent restatement (pure Python / numpy float64) of JoinCascador::Validate for a finished model:
    cascador.cpp:166-211, cart.cpp:392-404 (1-based heap), data.cpp:18-58 (round, clamp), btcart.cpp:407-424"""
    import math
    shape = mean + 0.0
    score, n = 0.0, 0
    for t in range(T):
        lbf = []
        for k in range(K):
            c = t * K + k
            node_idx = 1
            for _ in range(3):
                scale, l1, l2, o1x, o1y, o2x, o2y, th = nodes[c][node_idx - 1]
                assert scale == 0
                pts = []
                for s, o in ((shape[2 * l1], o1x), (shape[2 * l1 + 1], o1y), (shape[2 * l2], o2x), (shape[2 * l2 + 1], o2y)):
                    v = (s + o) * win
                    r = int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)      # C round(): half away from zero
                    pts.append(min(max(r, 0), win - 1))
                val = int(img[y + pts[1], x + pts[0]]) - int(img[y + pts[3], x + pts[2]])
                node_idx = 2 * node_idx if val <= th else 2 * node_idx + 1
            idx = node_idx - 8
            score += leaf[c][idx]
            score = (score - cart3[c][1]) / cart3[c][2]
            n += 1
            if score < cart3[c][0]:
                return False, score, n, None
            lbf.append(8 * k + idx)
        delta = np.zeros_like(shape)
        for row in lbf:
            delta = delta + w[t][row]
        shape = shape + delta
    return True, score, n, shape


def test_second_independent_restatement_agrees(ocpp, ocpp_shipped):
    """the C restatement against a from-scratch Python one that parses the model file itself: per-window carts
    evaluated and exit score (bit for bit), and the faces' landmarks"""
    raw = open(SHIPPED_F32, "rb").read()
    T, K, L = 5, 540, 27
    o = 28
    mean = np.frombuffer(raw, "<f4", 2 * L, o).astype(np.float64); o += 8 * L
    nodes, leaf, cart3, w = [], [], [], []
    for t in range(T):
        for k in range(K):
            nd = []
            for i in range(7):
                ints = np.frombuffer(raw, "<i4", 3, o); fl = np.frombuffer(raw, "<f4", 4, o + 12)
                th = int(np.frombuffer(raw, "<i4", 1, o + 28)[0])
                nd.append((int(ints[0]), int(ints[1]), int(ints[2]), float(fl[0]), float(fl[1]), float(fl[2]), float(fl[3]), th))
                o += 32
            nodes.append(nd)
            leaf.append(np.frombuffer(raw, "<f4", 8, o).astype(np.float64)); o += 32
            cart3.append(np.frombuffer(raw, "<f4", 3, o).astype(np.float64)); o += 12
        w.append(np.frombuffer(raw, "<f4", 8 * K * 2 * L, o).astype(np.float64).reshape(8 * K, 2 * L)); o += 4 * 8 * K * 2 * L
    assert o + 4 == len(raw)
    face = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "face_111x108.npy"))
    img = np.ascontiguousarray(face[10:60, 20:62])             # 42 x 50 crop: a handful of 20..39 px windows
    tn, ts = ocpp.trace(ocpp_shipped, img)
    wins = ocpp.levels(img.shape[1], img.shape[0])
    i = 0
    faces = []
    for win in wins:
        for y in range(0, img.shape[0] - win + 1, 5):
            for x in range(0, img.shape[1] - win + 1, 5):
                ok, score, n, shape = _py_validate(img, x, y, win, mean, nodes, leaf, cart3, w, T, K)
                assert n == tn[i], (i, n, tn[i])
                assert np.float64(score).view(np.uint64) == ts[i].view(np.uint64), (i, score, ts[i])
                if ok:
                    faces.append((x, y, win, shape))
                i += 1
    assert i == len(tn) and i >= 30
    rects, scores, shapes, _ = ocpp.detect(ocpp_shipped, img, nms=False)
    assert len(faces) == len(scores)
    for (x, y, win, shape), r, sh in zip(faces, rects, shapes):
        assert (x, y, win) == (r[0], r[1], r[2])
        want = shape.copy(); want[0::2] = x + shape[0::2] * win; want[1::2] = y + shape[1::2] * win
        np.testing.assert_array_equal(want.view(np.uint64), sh.view(np.uint64))


# ---- face.similarity_transform and a non-zero initial shift -------------------------------

@pytest.mark.parametrize("similarity,shift", [(True, 0.0), (False, 0.02), (True, 0.02)], ids=["similarity", "shift", "both"])
def test_similarity_transform_and_shift_pinned_to_reference(ocpp, ocpp_shipped, refcpp, wide_shipped, similarity, shift):
    """STParameter::Calc / Apply (data.cpp:64-126: offsets and the regressed delta go through the transform that maps
    the current shape onto the mean shape) and DataSet::RandomShape's shift (data.cpp:225-236), restatement vs the
    reference binary, bit for bit: Detect (raw + NMS) and Validate on every window.  The reference seeds the shift's RNG
    with the tick count for every window; the stand-in's getTickCount can be fixed, and the restatement is given the
    (x, y) the reference's own RNG then draws.  cv::norm inside Calc is the stand-in's (squares summed in index order):
    OpenCV's own accumulation order is not pinned by this."""
    hr = refcpp.load(wide_shipped)
    xy = refcpp.set_shift(shift, tick=123456789 if shift else 0)
    try:
        assert (xy == (0.0, 0.0)) == (shift == 0.0) and all(abs(v) <= shift for v in xy)
        ocpp.set_options(similarity=similarity, shift=xy)
        plain = None
        for name in ("faces", "facemix"):
            img = PIN_FRAMES[name]()
            for kw in (dict(), dict(minimum_size=30, step=7, scale=1.3, nms=False)):
                rr, rs, rsh, st = refcpp.detect(hr, img, similarity=similarity, **kw)
                orr, os_, osh, carts = ocpp.detect(ocpp_shipped, img, **kw)
                np.testing.assert_array_equal(rr, orr)
                _same64(rs, os_)
                _same64(rsh, osh)
                assert len(rs) > 0 or name != "faces"
            if name == "faces":
                plain = (rs, rsh)
        img = synth.facemix_frame(11, 200, 150)
        rn, rsc = refcpp.trace(hr, img, similarity=similarity)
        on, osc = ocpp.trace(ocpp_shipped, img)
        np.testing.assert_array_equal(rn, on)
        _same64(rsc, osc)
        assert (on > 540).any()
        # the options do change the numbers: against the plain run of the same frame some landmark differs
        ocpp.set_options()
        refcpp.set_shift(0.0, 0)
        _, s0, sh0, _ = ocpp.detect(ocpp_shipped, PIN_FRAMES["faces"](), minimum_size=30, step=7, scale=1.3, nms=False)
        assert len(s0) != len(plain[0]) or not np.array_equal(sh0, plain[1])
    finally:
        ocpp.set_options()
        refcpp.set_shift(0.0, 0)
        refcpp.release(hr)
