"""CUDA path of the reference's double-precision C++ detector (jdaB200JoinCascadorDetect: JoinCascador::Detect with
fddb.method = 1) against its CPU restatement oracle/jda_oracle_cpp.c, through the C ABI.

Bar: bit-exact.  Rects are integers; scores and landmarks are the same double operations in the same order, so
they are compared as raw 64-bit patterns.  The oracle is pinned bit for bit to the reference's own C++ detector built
into oracle/_ref_cpp/ (tests/test_oracle_cpp.py); test_against_the_reference_cpp_binary compares the CUDA path with
that binary directly.
"""
import numpy as np
import pytest

from jda_b200 import api, synth
from tests.conftest import SHIPPED_F32

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def _same(got, want):
    (r1, s1, p1), (r2, s2, p2) = got[:3], want[:3]
    assert r1.shape == r2.shape, (r1.shape, r2.shape)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(_bits(s1), _bits(s2))
    np.testing.assert_array_equal(_bits(p1), _bits(p2))


@pytest.fixture(scope="module")
def ocpp():
    from oracle import pyoracle
    return pyoracle.OracleCpp()


@pytest.fixture(scope="module")
def ocpp_shipped(ocpp):
    h = ocpp.load(SHIPPED_F32, double=False)
    yield h
    ocpp.release(h)


@pytest.fixture(scope="module")
def casc():
    c = api.Cascador(SHIPPED_F32, double=False)
    yield c
    c.close()


def test_known_answer_config_json_settings(casc, ocpp, ocpp_shipped):
    img = synth.face_canvas()
    got = casc.detect_cpp(img)
    assert got[0].tolist() == [[400, 290, 112, 112], [55, 40, 230, 230]]
    _same(got, ocpp.detect(ocpp_shipped, img))
    st = casc.last_stats
    assert st["windows"] == 140215 and st["scan_launches"] == 1 and st["raw_hits"] == 365
    assert st["stage0_survivors"] >= 365                      # the prefilter passes a superset of the true survivors


def test_against_the_reference_cpp_binary(casc, tmp_path):
    """the CUDA path vs the reference's own JoinCascador::Detect (src/jda/cascador.cpp compiled into oracle/_ref_cpp),
    no restatement in between"""
    import os
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_CPP_SO):
        pytest.skip("oracle/_ref_cpp/libjda_ref_cpp.so not built")
    ref = pyoracle.RefCpp()
    hr = ref.load(synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model")))
    assert hr
    for img, kw in ((synth.face_canvas(), dict()), (synth.facemix_frame(3, 320, 240), dict(nms=False)),
                    (synth.facemix_frame(7, *synth.fddb_shape(7)), dict(minimum_size=30, step=7, scale=1.3, overlap=0.5))):
        want = ref.detect(hr, img, **kw)
        _same(casc.detect_cpp(img, **kw), want)
        assert casc.last_stats["windows"] == want[3]["patch_n"] and casc.last_stats["raw_hits"] == want[3]["face_patch_n"]
    tn, ts = casc.trace_cpp(synth.noise_frame(2, 90, 70))
    rn, rs = ref.trace(hr, synth.noise_frame(2, 90, 70))
    np.testing.assert_array_equal(tn, rn)
    np.testing.assert_array_equal(_bits(ts), _bits(rs))
    ref.release(hr)


@pytest.mark.parametrize("frame", ["faces", "noise", "blur", "facemix", "odd"])
def test_raw_hits_and_nms_match_oracle(casc, ocpp, ocpp_shipped, frame):
    img = {"faces": synth.face_canvas, "noise": lambda: synth.noise_frame(4), "blur": lambda: synth.blur_frame(2),
           "facemix": lambda: synth.facemix_frame(11), "odd": lambda: synth.facemix_frame(12, 333, 251)}[frame]()
    for kw in (dict(nms=False), dict(), dict(minimum_size=30, step=3, scale=1.3, overlap=0.5)):
        _same(casc.detect_cpp(img, **kw), ocpp.detect(ocpp_shipped, img, **kw))


def test_per_window_trace_double_kernel(casc, ocpp, ocpp_shipped):
    """every window through k4_cascade_f64: carts evaluated (Validate's n) and the exit score, bit for bit"""
    for img in (synth.facemix_frame(21, 200, 150), synth.face_canvas()):
        tn, ts = casc.trace_cpp(img)
        on, os_ = ocpp.trace(ocpp_shipped, img)
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))


def test_prefilter_equals_dense_double_evaluation(casc, ocpp, ocpp_shipped):
    """the float32 scan is only a filter: with it (default) and without it (every window in double) the answers
    are the same bits"""
    frames = synth.make_frames("facemix", 5, 320, 240, seed0=40)
    a = casc.detect_cpp(frames, nms=False)
    sa = dict(casc.last_stats)
    b = casc.detect_cpp(frames, nms=False, flags=api.NO_STAGE0_SCAN)
    sb = dict(casc.last_stats)
    assert sa["scan_launches"] == 1 and sb["scan_launches"] == 0 and sa["raw_hits"] == sb["raw_hits"]
    for x, y, f in zip(a, b, frames):
        _same(x, y)
        _same(x, ocpp.detect(ocpp_shipped, f, nms=False))


def test_batch_and_chunked_batch(casc, ocpp, ocpp_shipped):
    frames = synth.make_frames("facemix", 130, 160, 120, seed0=70)      # >= 128 frames: chunked copies + scans
    frames[5, 6:114, 20:131] = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "face_111x108.npy"))
    got = casc.detect_cpp(frames)
    assert casc.last_stats["scan_launches"] == 4
    assert casc.last_stats["windows"] == 130 * api.count_windows_cpp(160, 120)
    assert sum(len(g[1]) for g in got) >= 1
    for i in (0, 5, 64, 129):
        _same(got[i], ocpp.detect(ocpp_shipped, frames[i]))


SYN = [dict(seed=1, mode="passall"), dict(seed=2, mode="reject"),
       dict(seed=5, mode="reject", norm_every=7)]      # > 32 normalised carts: no prefilter, dense double evaluation


@pytest.mark.parametrize("cfg", SYN, ids=lambda c: "seed%d" % c["seed"])
def test_synthetic_double_models(ocpp, tmp_path, cfg):
    path = synth.write_model(str(tmp_path / "syn.model"), **cfg)
    c = api.Cascador(path, double=True)
    h = ocpp.load(path, True)
    assert (c.filter_margins_cpp() is None) == (cfg.get("norm_every") == 7)
    for img in (synth.blur_frame(9, 96, 80), synth.noise_frame(5, 70, 61)):
        for kw in (dict(nms=False), dict(minimum_size=24, step=4, scale=1.25, overlap=0.3)):
            _same(c.detect_cpp(img, **kw), ocpp.detect(h, img, **kw))
        tn, ts = c.trace_cpp(img)
        on, os_ = ocpp.trace(h, img)
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
    c.close(); ocpp.release(h)


@pytest.mark.parametrize("dims", [dict(T=2, K=33, L=5), dict(T=3, K=96, L=40), dict(T=2, K=1000, L=8)],
                         ids=lambda d: "T%d_K%d_L%d" % (d["T"], d["K"], d["L"]))
def test_model_dimensions_from_the_header(ocpp, tmp_path, dims):
    path = synth.write_model(str(tmp_path / "dims.model"), seed=11 + dims["K"], mode="reject", norm_every=270 if dims["K"] > 900 else 10, **dims)
    c = api.Cascador(path, double=True)
    h = ocpp.load(path, True)
    for img in (synth.blur_frame(9, 96, 80), synth.facemix_frame(3, 200, 150)):
        _same(c.detect_cpp(img, nms=False), ocpp.detect(h, img, nms=False))
        _same(c.detect_cpp(img), ocpp.detect(h, img))
    tn, ts = c.trace_cpp(synth.noise_frame(5, 70, 61))
    on, os_ = ocpp.trace(h, synth.noise_frame(5, 70, 61))
    np.testing.assert_array_equal(tn, on)
    np.testing.assert_array_equal(_bits(ts), _bits(os_))
    c.close(); ocpp.release(h)


def test_training_snapshot_header(ocpp, tmp_path):
    """a model whose header stops inside a stage (cascador.cpp:199-209): full stages, then carts 0..cart, no regression"""
    q = synth.write_model(str(tmp_path / "rej.model"), seed=2, mode="reject")
    for stage, cart in ((2, 17), (0, 40), (0, -1), (4, 539)):
        b = bytearray(open(q, "rb").read())
        b[20:24] = stage.to_bytes(4, "little"); b[24:28] = cart.to_bytes(4, "little", signed=True)
        p = tmp_path / ("snap_%d_%d.model" % (stage, cart))
        p.write_bytes(bytes(b))
        c = api.Cascador(str(p), double=True)
        h = ocpp.load(str(p), True)
        img = synth.blur_frame(13, 90, 77)
        _same(c.detect_cpp(img, nms=False), ocpp.detect(h, img, nms=False))
        tn, ts = c.trace_cpp(img)
        on, os_ = ocpp.trace(h, img)
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        c.close(); ocpp.release(h)


def test_refusals_and_edges(casc, ocpp, ocpp_shipped, tmp_path):
    p = synth.write_model(str(tmp_path / "scaled.model"), seed=3, scales=(0, 1, 2), coord_max=0.45)
    c = api.Cascador(p, double=True)
    with pytest.raises(RuntimeError, match="scale != 0"):
        c.detect_cpp(synth.noise_frame(0, 64, 48))
    c.close()
    with pytest.raises(RuntimeError):
        casc.detect_cpp(synth.noise_frame(0, 64, 48), step=0)
    with pytest.raises(RuntimeError):
        casc.detect_cpp(synth.noise_frame(0, 64, 48), overlap=1.0)
    tiny = synth.noise_frame(1, 19, 40)                      # smaller than fddb.minimum_size: no window
    got = casc.detect_cpp(tiny)
    assert len(got[1]) == 0 and casc.last_stats["windows"] == 0
    one = synth.blur_frame(2, 20, 20)                        # exactly one window
    _same(casc.detect_cpp(one, nms=False), ocpp.detect(ocpp_shipped, one, nms=False))
    assert len(casc.detect_cpp(synth.noise_frame(1), scale=1.0)[1]) == 0     # endless loop in the reference: no detections
    # the float C path on the same handle is untouched by the double path's state
    img = synth.face_canvas()
    assert casc.detect(img)[0].tolist() == [[396, 308, 110], [63, 21, 213]]


def test_cli_fddb_runner_uses_the_cpp_detector(tmp_path, ocpp, ocpp_shipped):
    """src/test.cpp:73-235 (the FDDB runner) calls JoinCascador::Detect and writes `x y w h score` per face"""
    from jda_b200.__main__ import main
    img, small = synth.face_canvas(), synth.facemix_frame(8, 300, 200)
    np.save(tmp_path / "a.npy", img)
    np.save(tmp_path / "b.npy", small)
    out = tmp_path / "fold-01-out.txt"
    assert main(["detect", SHIPPED_F32, str(tmp_path / "a.npy"), str(tmp_path / "b.npy"), "--float", "--cpp",
                 "--fddb-out", str(out)]) == 0
    lines = out.read_text().split("\n")
    r, s, _, _ = ocpp.detect(ocpp_shipped, img)
    assert lines[0] == str(tmp_path / "a") and lines[1] == "2"
    assert lines[2] == "%d %d %d %d %f" % (r[0][0], r[0][1], r[0][2], r[0][3], s[0])
    r2, s2, _, _ = ocpp.detect(ocpp_shipped, small)
    assert lines[4] == str(tmp_path / "b") and lines[5] == str(len(s2))


def test_fddb_runner_layout(tmp_path, ocpp, ocpp_shipped):
    """python -m jda_b200 fddb: folds in, result/fold-XX-out.txt out (src/test.cpp:96-224), both detectors"""
    cv2 = pytest.importorskip("cv2")
    from jda_b200.__main__ import main
    root = tmp_path / "fddb"
    (root / "FDDB-folds").mkdir(parents=True)
    (root / "images" / "2002" / "07").mkdir(parents=True)
    imgs = {"2002/07/img_1": synth.face_canvas(), "2002/07/img_2": synth.facemix_frame(5, 410, 300)}
    for k, v in imgs.items():
        assert cv2.imwrite(str(root / "images" / (k + ".jpg")), cv2.cvtColor(v, cv2.COLOR_GRAY2BGR), [cv2.IMWRITE_JPEG_QUALITY, 97])
    (root / "FDDB-folds" / "FDDB-fold-01.txt").write_text("2002/07/img_1\n2002/07/missing\n2002/07/img_2\n")
    assert main(["fddb", SHIPPED_F32, str(root), "--float", "--folds", "1"]) == 0
    lines = (root / "result" / "fold-01-out.txt").read_text().split("\n")
    gray = cv2.cvtColor(cv2.imread(str(root / "images" / "2002/07/img_1.jpg")), cv2.COLOR_BGR2GRAY)   # what the runner saw
    r, s, _, _ = ocpp.detect(ocpp_shipped, gray)
    assert lines[0] == "2002/07/img_1" and lines[1] == str(len(s)) and len(s) >= 2
    assert lines[2] == "%d %d %d %d %f" % (r[0][0], r[0][1], r[0][2], r[0][3], s[0])
    assert lines[2 + len(s)] == "2002/07/img_2"                     # the unreadable image was skipped
    assert main(["fddb", SHIPPED_F32, str(root), "--float", "--folds", "1", "--c-api"]) == 0
    lines = (root / "result" / "fold-01-out.txt").read_text().split("\n")
    assert lines[0] == "2002/07/img_1" and int(lines[1]) >= 2 and len(lines[2].split()) == 5


# ---- face.similarity_transform and the initial shift ------------------------------------

@pytest.mark.parametrize("similarity,shift", [(True, (0.0, 0.0)), (False, (0.013, -0.0071)), (True, (-0.02, 0.0175))],
                         ids=["similarity", "shift", "both"])
def test_similarity_transform_and_initial_shift(casc, ocpp, ocpp_shipped, similarity, shift):
    """STParameter::Calc / Apply per stage and mean_shape + (x, y) as the initial shape, CUDA path vs the restatement
    (which tests/test_oracle_cpp.py pins to the reference binary): raw hits, NMS, a batch, the per-window trace.
    With a zero shift the float32 prefilter stays on (the stage-0 transform of shape == mean shape is the identity, bit
    for bit); a shifted initial shape sends every window through the double kernel."""
    ocpp.set_options(similarity=similarity, shift=shift)
    try:
        kw = dict(similarity=similarity, shift=shift)
        for img, okw in ((synth.face_canvas(), dict()), (synth.facemix_frame(11), dict(nms=False)),
                         (synth.facemix_frame(12, 333, 251), dict(minimum_size=30, step=3, scale=1.3, overlap=0.5))):
            _same(casc.detect_cpp(img, **okw, **kw), ocpp.detect(ocpp_shipped, img, **okw))
            assert casc.last_stats["scan_launches"] == (1 if shift == (0.0, 0.0) else 0)
        frames = synth.make_frames("facemix", 6, 320, 240, seed0=70)
        for got, f in zip(casc.detect_cpp(frames, nms=False, **kw), frames):
            _same(got, ocpp.detect(ocpp_shipped, f, nms=False))
        img = synth.facemix_frame(21, 200, 150)
        tn, ts = casc.trace_cpp(img, **kw)
        on, os_ = ocpp.trace(ocpp_shipped, img)
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        assert (on > 1080).any()                                  # windows that went through two regressions
    finally:
        ocpp.set_options()
    # and the options do change the answer
    a = casc.detect_cpp(synth.face_canvas(), nms=False)
    b = casc.detect_cpp(synth.face_canvas(), nms=False, similarity=similarity, shift=shift)
    assert len(a[1]) != len(b[1]) or not np.array_equal(a[2], b[2])


def test_similarity_and_shift_against_the_reference_cpp_binary(casc, tmp_path):
    """the same, CUDA path vs the reference's own binary with RandomShape's seed fixed (no restatement in between)"""
    import os
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_CPP_SO):
        pytest.skip("oracle/_ref_cpp/libjda_ref_cpp.so not built")
    ref = pyoracle.RefCpp()
    hr = ref.load(synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model")))
    try:
        xy = ref.set_shift(0.02, tick=987654321)
        assert xy != (0.0, 0.0)
        for img, kw in ((synth.face_canvas(), dict()), (synth.facemix_frame(3, 320, 240), dict(nms=False))):
            _same(casc.detect_cpp(img, similarity=True, shift=xy, **kw), ref.detect(hr, img, similarity=True, **kw))
        xy = ref.set_shift(0.0, 0)
        _same(casc.detect_cpp(synth.face_canvas(), similarity=True), ref.detect(hr, synth.face_canvas(), similarity=True))
    finally:
        ref.set_shift(0.0, 0)
        ref.release(hr)
