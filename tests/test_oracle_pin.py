"""Pins the oracle restatement (oracle/jda_oracle.c) to the reference.

(1) against committed outputs of the reference's own jdaDetect (tests/golden/ref_outputs.npz,
    made by tests/golden/make_fixtures.py);
(2) against the reference library itself (oracle/_ref/libjda_ref.so) on synthetic models and
    frames, whenever that library is present.
Bit-exact throughout: the restatement is the same float32 arithmetic in the same order.
"""
import os

import numpy as np
import pytest

from jda_b200 import synth
from tests.conftest import SHIPPED_F32, REF_SHIPPED_F64
from tests.golden.make_fixtures import CASES


def _same(a, b):
    (b1, s1, p1), (b2, s2, p2) = a, b
    assert b1.shape == b2.shape, (b1.shape, b2.shape)
    np.testing.assert_array_equal(b1, b2)
    np.testing.assert_array_equal(s1.view(np.uint32), s2.view(np.uint32))
    np.testing.assert_array_equal(p1.view(np.uint32), p2.view(np.uint32))


@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_oracle_matches_reference_golden(oracle, oracle_shipped, gold, case):
    name, mk, kw = next(c for c in CASES if c[0] == case)
    frame = mk(synth)
    crc = gold[name + "/crc"]
    assert [int(frame.astype(np.uint64).sum()), frame.shape[1], frame.shape[0]] == crc.tolist(), \
        "synthetic frame generator drifted from the one the golden vectors were made with"
    got = oracle.detect(oracle_shipped, frame, **kw)
    _same(got, (gold[name + "/boxes"], gold[name + "/scores"], gold[name + "/shapes"]))


def test_known_answer_face_canvas(oracle, oracle_shipped):
    # SURVEY.md section 4 / BASELINE.md: the reference's only shipped real-face input
    boxes, scores, shapes = oracle.detect(oracle_shipped, synth.face_canvas(), th=0.0)
    assert boxes.tolist() == [[396, 308, 110], [63, 21, 213]]
    np.testing.assert_allclose(scores, [1.612964, 1.619511], rtol=0, atol=5e-7)
    assert shapes.shape == (2, 54)
    for b, s in zip(boxes, shapes):
        assert (s[0::2] > b[0] - 0.3 * b[2]).all() and (s[0::2] < b[0] + 1.3 * b[2]).all()


def test_window_counts(oracle):
    # SURVEY.md section 8 header
    assert oracle.levels(640, 480) == [24, 30, 37, 46, 57, 71, 88, 110, 137, 171, 213, 266, 332, 415]
    assert oracle.count_windows(640, 480) == 169706
    assert oracle.count_windows(640, 480, max_size=192) == 169236
    assert oracle.count_windows(640, 480, scale=1.2) == 245217
    assert oracle.count_windows(640, 480, min_size=40) == 38245
    assert oracle.count_windows(1920, 1080, max_size=768) == 1245202
    assert oracle.count_windows(1920, 1080) == 1245268
    assert oracle.count_windows(23, 400) == 0
    assert oracle.count_windows(640, 480, scale=1.0) == 0   # reference would spin forever


def test_work_statistics(oracle, oracle_shipped):
    # BASELINE.md section 2 (measured on the reference during the survey)
    _, _, _, st = oracle.detect_raw(oracle_shipped, synth.noise_frame(0))
    assert st["windows"] == 169706
    assert abs(st["carts"] / st["windows"] - 99.36) < 0.01
    assert st["stage_survivors"][:5] == [46, 0, 0, 0, 0]
    b, s, p, st = oracle.detect_raw(oracle_shipped, synth.face_canvas(), use_th=False)
    assert st["stage_survivors"][:5] == [181, 60, 40, 39, 37]
    assert len(s) == 37 and st["ub_reads"] == 0
    assert abs(st["carts"] / st["windows"] - 17.93) < 0.01


def test_trace_consistent_with_raw(oracle, oracle_shipped):
    img = synth.face_canvas()
    tn, ts, lv = oracle.trace(oracle_shipped, img, leaf_range=(0, 2000))
    b, s, p, st = oracle.detect_raw(oracle_shipped, img, use_th=False)
    assert tn.sum() == st["carts"] and (tn == 2700).sum() == len(s)
    np.testing.assert_array_equal(ts[tn == 2700].view(np.uint32), s.view(np.uint32))
    # evaluated carts have a leaf in 0..7, the rest stay 255
    for i in (0, 17, 1999):
        assert (lv[i, :tn[i]] < 8).all() and (lv[i, tn[i]:] == 255).all()


def test_serialiser_bytes(oracle, oracle_shipped, tmp_path):
    # golden file was written by the REFERENCE's jdaCascadorSerializeTo
    out = tmp_path / "rt.model"
    assert oracle.save_f32(oracle_shipped, str(out)) == 0
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()


def test_widened_double_file_loads_identically(oracle, oracle_shipped, tmp_path):
    wide = synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model"))
    assert os.path.getsize(wide) == 10476464
    h = oracle.load(wide, double=True)
    out = tmp_path / "rt.model"
    oracle.save_f32(h, str(out))
    oracle.release(h)
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF_SHIPPED_F64), reason="/root/reference not present")
def test_shipped_double_model_matches_golden_f32(oracle, tmp_path):
    h = oracle.load(REF_SHIPPED_F64, double=True)
    out = tmp_path / "rt.model"
    oracle.save_f32(h, str(out))
    oracle.release(h)
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()


# ---- directly against the reference library on synthetic models ------------------------------

SYN = [
    dict(seed=1, mode="passall", scales=(0,)),
    dict(seed=2, mode="reject", scales=(0,)),
    dict(seed=3, mode="reject", scales=(0, 1, 2), coord_max=0.45),
    dict(seed=4, mode="passall", scales=(0, 1, 2), coord_max=0.45, double=False),
    # stage 0 on the o plane only, h / q nodes from stage 1 on: the model shape that takes the LUT scan AND the planes
    dict(seed=6, mode="reject", scales=(0, 1, 2), coord_max=0.45, scales_by_stage={0: (0,)}),
    dict(seed=7, mode="passall", scales=(0, 1, 2), coord_max=0.45, scales_by_stage={0: (0,)}),
]


@pytest.mark.parametrize("cfg", SYN, ids=lambda c: "seed%d-%s-%s" % (c["seed"], c["mode"], len(c["scales"])))
def test_oracle_vs_reference_library_synthetic(oracle, reflib, tmp_path, cfg):
    cfg = dict(cfg)
    dbl = cfg.pop("double", True)
    path = synth.write_model(str(tmp_path / "syn.model"), double=dbl, **cfg)
    ho, hr = oracle.load(path, dbl), reflib.load(path, dbl)
    assert ho and hr
    # serialisers agree byte for byte
    a, b = tmp_path / "a.model", tmp_path / "b.model"
    oracle.save_f32(ho, str(a)); reflib.save_f32(hr, str(b))
    assert a.read_bytes() == b.read_bytes()
    frames = [synth.blur_frame(9, 96, 80), synth.noise_frame(5, 70, 61)]
    if cfg["mode"] == "reject":
        frames.append(synth.blur_frame(10, 320, 240))
    for img in frames:
        for kw in (dict(scale=1.25, min_size=24, max_size=-1, th=-1e30),
                   dict(scale=1.3, min_size=30, max_size=60, th=0.5)):
            got, want = oracle.detect(ho, img, **kw), reflib.detect(hr, img, **kw)
            _same(got, want)
            if cfg["mode"] == "passall" and kw["th"] < -1e29:
                assert len(got[1]) >= 1
        _, _, _, st = oracle.detect_raw(ho, img)
        assert st["ub_reads"] == 0, "synthetic model strayed into the reference's UB region"
    oracle.release(ho); reflib.release(hr)


def test_oracle_vs_reference_library_shipped(oracle, reflib, oracle_shipped):
    hr = reflib.load(SHIPPED_F32, double=False)
    for img, kw in [(synth.facemix_frame(21), dict(th=-2.0)),
                    (synth.facemix_frame(22, 450, 333), dict(scale=1.15, min_size=28, max_size=200, th=0.0))]:
        _same(oracle.detect(oracle_shipped, img, **kw), reflib.detect(hr, img, **kw))
    reflib.release(hr)


def test_nms_passall_many_hits(oracle, reflib, tmp_path):
    """hundreds of overlapping raw hits through the reference's NMS vs ours."""
    path = synth.write_model(str(tmp_path / "p.model"), seed=8, mode="passall")
    ho, hr = oracle.load(path, True), reflib.load(path, True)
    img = synth.blur_frame(3, 128, 100)
    got, want = oracle.detect(ho, img, th=-1e30), reflib.detect(hr, img, th=-1e30)
    raw = oracle.detect_raw(ho, img, th=-1e30)
    assert len(raw[1]) == oracle.count_windows(128, 100) and len(want[1]) < len(raw[1])
    _same(got, want)
    oracle.release(ho); reflib.release(hr)


def test_oracle_cart_granular_truncation_properties(oracle, oracle_shipped):
    """k_limit (Validate's unfinished stage, src/jda/cascador.cpp:199-209) has no reference binary behind it on the
    float path, so it is pinned through properties against the reference-pinned whole-stage cascade: stopping after
    ALL K carts of stage t passes exactly the windows that t + 1 whole stages pass, with the same scores, and their
    shapes are the shapes after t whole stages (no regression follows an unfinished stage); carts evaluated never
    exceed t*K + k and windows that die earlier die at the same cart with the same score."""
    from jda_b200 import synth
    img = synth.face_canvas()
    K = 540
    for t in (0, 1, 3):
        pb, ps, psh, _ = oracle.detect_raw(oracle_shipped, img, t_limit=t, k_limit=K, use_th=False)
        fb, fs, fsh, _ = oracle.detect_raw(oracle_shipped, img, t_limit=t + 1, use_th=False)
        np.testing.assert_array_equal(pb, fb)
        np.testing.assert_array_equal(ps.view(np.uint32), fs.view(np.uint32))
        assert len(ps) > 0 and not np.array_equal(psh, fsh)
        if t > 0:
            tb, _, tsh, _ = oracle.detect_raw(oracle_shipped, img, t_limit=t, use_th=False)
            keys = {tuple(b): i for i, b in enumerate(tb.tolist())}
            idx = [keys[tuple(b)] for b in pb.tolist()]
            np.testing.assert_array_equal(psh.view(np.uint32), tsh[idx].view(np.uint32))
    full_n, full_s, _ = oracle.trace(oracle_shipped, img)
    for t, k in ((0, 18), (2, 101)):
        n, s, _ = oracle.trace(oracle_shipped, img, t_limit=t, k_limit=k)
        cap = t * K + k
        np.testing.assert_array_equal(n, np.minimum(full_n, cap))
        early = full_n < cap
        np.testing.assert_array_equal(s[early].view(np.uint32), full_s[early].view(np.uint32))
        assert (n == cap).sum() > 0
