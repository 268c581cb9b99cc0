"""Parity of the CUDA path against the oracle and the reference's golden outputs.

Everything goes through the C ABI (jda_b200.api is a ctypes shim over it).  Bar: bit-exact --
tree traversal, leaf indices, reject positions are integer work; scores and landmarks are the
same float32 operations in the same order, so they are compared as raw bits (tolerance 0, which is
inside north_star's 1e-6 relative).
"""
import os

import numpy as np
import pytest

from jda_b200 import api, synth
from tests.conftest import SHIPPED_F32
from tests.golden.make_fixtures import CASES

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(got, want):
    (b1, s1, p1), (b2, s2, p2) = got, want
    assert b1.shape == b2.shape, (b1.shape, b2.shape)
    np.testing.assert_array_equal(b1, b2)
    np.testing.assert_array_equal(_bits(s1), _bits(s2))
    np.testing.assert_array_equal(_bits(p1), _bits(p2))


@pytest.fixture(scope="module")
def casc():
    c = api.Cascador(SHIPPED_F32, double=False)
    yield c
    c.close()


def _with_env(*names, **values):
    """A handle created under A/B knobs / test hooks (a handle reads them from the environment once, when it is created)."""
    env = {n: "1" for n in names}
    env.update({k: str(v) for k, v in values.items()})
    os.environ.update(env)
    try:
        return api.Cascador(SHIPPED_F32, double=False)
    finally:
        for k in env:
            del os.environ[k]


@pytest.fixture(scope="module")
def casc_stages():
    """throughput tile plan and stages >= 1 through k3_walk / k3_regress / k3_emit whatever the batch size (by default
    batches under 2.1e6 candidate windows take the latency plan, and batches under 6e7 -- about 350 VGA frames -- take
    k3_cascade: ten small launches are pure latency on a short survivor list)"""
    c = _with_env(JDA_B200_STAGE_MIN_WINDOWS=0, JDA_B200_FORCE_PLAN="throughput")
    yield c
    c.close()


# ---- the reference's own outputs -------------------------------------------------------------

@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_golden_reference_outputs(casc, gold, case):
    name, mk, kw = next(c for c in CASES if c[0] == case)
    got = casc.detect(mk(synth), **kw)
    _same(got, (gold[name + "/boxes"], gold[name + "/scores"], gold[name + "/shapes"]))


def test_known_answer_and_landmark_rmse(casc, oracle, oracle_shipped):
    img = synth.face_canvas()
    got = casc.detect(img)
    assert got[0].tolist() == [[396, 308, 110], [63, 21, 213]]
    want = oracle.detect(oracle_shipped, img)
    rmse = float(np.sqrt(np.mean((got[2].astype(np.float64) - want[2]) ** 2)))
    assert rmse == 0.0
    _same(got, want)


# ---- per-window trace: reject position, exit score, every evaluated leaf ------------------------

TRACE_FRAMES = {
    "noise": lambda: synth.noise_frame(0),
    "blur": lambda: synth.blur_frame(1),
    "faces": lambda: synth.face_canvas(),
    "odd_size": lambda: synth.facemix_frame(4, 451, 333),
    "narrow": lambda: synth.facemix_frame(6, 60, 300),      # few windows per row: 8-wide tiles
    "wide": lambda: synth.facemix_frame(8, 400, 64),
}


@pytest.mark.parametrize("flags", [0, api.NO_TMA], ids=["tma", "plain_loads"])
@pytest.mark.parametrize("frame", list(TRACE_FRAMES))
def test_trace_matches_oracle(casc, oracle, oracle_shipped, frame, flags):
    img = TRACE_FRAMES[frame]()
    nwin = api.count_windows(img.shape[1], img.shape[0])
    # leaves for three slices of the scan: first windows, the middle, the coarse levels at the end
    for rng in [(0, 3000), (nwin // 2, nwin // 2 + 3000), (nwin - 3000, nwin)]:
        tn, ts, lv = casc.trace(img, flags=flags, leaf_range=rng)
        on, os_, olv = oracle.trace(oracle_shipped, img, leaf_range=rng)
        np.testing.assert_array_equal(tn, on)           # carts evaluated = reject (stage, cart)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        np.testing.assert_array_equal(lv, olv)


@pytest.mark.parametrize("flags", [0, api.NO_TMA], ids=["tma", "plain_loads"])
@pytest.mark.parametrize("frame", ["noise", "faces", "odd_size", "narrow", "wide"])
def test_trace_throughput_plan(oracle, oracle_shipped, frame, flags):
    """The batch (throughput) tile plan on one frame: pooled tiles, 512-window lists, global-memory levels -- every
    window's reject cart, exit score and leaves.  JDA_B200_FORCE_PLAN is a test hook: a one-frame call (the trace entry
    point) would otherwise only ever take the latency plan."""
    c = _with_env(JDA_B200_FORCE_PLAN="throughput", JDA_B200_STAGE_MIN_WINDOWS=0)   # (and the traced stage kernels)
    img = TRACE_FRAMES[frame]()
    nwin = api.count_windows(img.shape[1], img.shape[0])
    for rng in [(max(nwin - 9000, 0), nwin), (nwin // 2, nwin // 2 + 2000), (0, 2000)]:
        tn, ts, lv = c.trace(img, flags=flags, leaf_range=rng)
        on, os_, olv = oracle.trace(oracle_shipped, img, leaf_range=rng)
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        np.testing.assert_array_equal(lv, olv)
    _same(c.detect(img, th=-1.0), oracle.detect(oracle_shipped, img, th=-1.0))
    c.close()


def test_trace_generic_kernel_only(casc, oracle, oracle_shipped):
    """stage-0 scan switched off: every window through k3_cascade."""
    img = synth.face_canvas()[:240, :320].copy()
    tn, ts, lv = casc.trace(img, flags=api.NO_STAGE0_SCAN, leaf_range=(0, 4000))
    on, os_, olv = oracle.trace(oracle_shipped, img, leaf_range=(0, 4000))
    np.testing.assert_array_equal(tn, on)
    np.testing.assert_array_equal(_bits(ts), _bits(os_))
    np.testing.assert_array_equal(lv, olv)


@pytest.mark.parametrize("stragglers", ["1", "0"])
@pytest.mark.parametrize("nw", ["1", "2", "4"])
def test_scan_kernel_variants(oracle, oracle_shipped, nw, stragglers):
    """windows per lane (ILP width) and straggler mode on/off are scheduling choices only."""
    os.environ["JDA_B200_NW"] = nw
    os.environ["JDA_B200_STRAGGLERS"] = stragglers
    try:
        c = api.Cascador(SHIPPED_F32, double=False)
        img = synth.facemix_frame(9)
        _same(c.detect(img, th=-1.0), oracle.detect(oracle_shipped, img, th=-1.0))
        raw = c.detect_batch(img[None], th=0.0, flags=api.RAW_HITS | api.NO_FINAL_TH)[0]
        ob, osc, osh, st = oracle.detect_raw(oracle_shipped, img, use_th=False)
        _same(raw, (ob, osc, osh))
        assert c.last_stats["stage0_survivors"] == st["stage_survivors"][0]
        for im in (img, synth.noise_frame(3, 200, 150)):
            nwin = api.count_windows(im.shape[1], im.shape[0])
            rng = (max(0, nwin - 6000), nwin)
            tn, ts, lv = c.trace(im, leaf_range=rng)
            on, os_, olv = oracle.trace(oracle_shipped, im, leaf_range=rng)
            np.testing.assert_array_equal(tn, on)
            np.testing.assert_array_equal(_bits(ts), _bits(os_))
            np.testing.assert_array_equal(lv, olv)
        c.close()
    finally:
        del os.environ["JDA_B200_NW"]
        del os.environ["JDA_B200_STRAGGLERS"]


def test_tile_origins_not_multiple_of_16(oracle, oracle_shipped):
    """Tiles whose x origin is not 16-byte aligned (8-wide tiles at step 5: origin 40*tx): the TMA box starts
    at the aligned address below and the windows are addressed with the remainder."""
    os.environ["JDA_B200_MIN_TILE_WINDOWS"] = "32"
    try:
        assert any(p["smem"] and (p["tw"] * p["step"]) % 16 for p in api.describe_plan(640, 480, latency=True))
        c = api.Cascador(SHIPPED_F32, double=False)
        img = synth.face_canvas()
        for flags in (0, api.NO_TMA):
            got = c.detect_batch(img[None], th=0.0, flags=flags)[0]
            _same(got, oracle.detect(oracle_shipped, img))
        nwin = api.count_windows(640, 480)
        tn, ts, lv = c.trace(img, leaf_range=(nwin - 40000, nwin - 36000))
        on, os_, olv = oracle.trace(oracle_shipped, img, leaf_range=(nwin - 40000, nwin - 36000))
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        np.testing.assert_array_equal(lv, olv)
        c.close()
    finally:
        del os.environ["JDA_B200_MIN_TILE_WINDOWS"]


@pytest.mark.parametrize("span", ["1", "2", "4", "12"])
def test_pooled_tile_buffers(oracle, oracle_shipped, span):
    """coarse levels served from tiles that span several warps' buffers, rows split across the group"""
    os.environ["JDA_B200_MAX_SPAN"] = span
    try:
        assert span == "1" or any(p["span"] > 1 for p in api.describe_plan(640, 480, latency=True))
        c = api.Cascador(SHIPPED_F32, double=False)
        for img in (synth.face_canvas(), synth.noise_frame(4)):
            nwin = api.count_windows(640, 480)
            for flags in (0, api.NO_TMA):
                tn, ts, lv = c.trace(img, flags=flags, leaf_range=(nwin - 12000, nwin - 8000))
                on, os_, olv = oracle.trace(oracle_shipped, img, leaf_range=(nwin - 12000, nwin - 8000))
                np.testing.assert_array_equal(tn, on)
                np.testing.assert_array_equal(_bits(ts), _bits(os_))
                np.testing.assert_array_equal(lv, olv)
        _same(c.detect(synth.face_canvas()), oracle.detect(oracle_shipped, synth.face_canvas()))
        c.close()
    finally:
        del os.environ["JDA_B200_MAX_SPAN"]


# ---- synthetic models: deep survivors, normalised scores, scaled (h/q) nodes --------------------

SYN = [
    dict(seed=1, mode="passall", scales=(0,)),
    dict(seed=2, mode="reject", scales=(0,)),
    dict(seed=3, mode="reject", scales=(0, 1, 2), coord_max=0.45),
    dict(seed=4, mode="passall", scales=(0, 1, 2), coord_max=0.45),
    dict(seed=5, mode="reject", scales=(0,), norm_every=7),      # > 32 normalised carts: generic path
    # stage 0 all scale 0 (LUT scan), later stages sample h / q: k2_scan -> k3_stage0 -> k3_cascade reading the planes
    dict(seed=6, mode="reject", scales=(0, 1, 2), coord_max=0.45, scales_by_stage={0: (0,)}),
    dict(seed=7, mode="passall", scales=(0, 1, 2), coord_max=0.45, scales_by_stage={0: (0,)}),
]


@pytest.mark.parametrize("cfg", SYN, ids=lambda c: "seed%d" % c["seed"])
def test_synthetic_models(oracle, tmp_path, cfg):
    path = synth.write_model(str(tmp_path / "syn.model"), **cfg)
    c = api.Cascador(path, double=True)
    ho = oracle.load(path, True)
    frames = [synth.blur_frame(9, 96, 80), synth.noise_frame(5, 70, 61)]
    if cfg["mode"] == "reject":
        frames.append(synth.blur_frame(10, 320, 240))
    for img in frames:
        for kw in (dict(scale=1.25, min_size=24, max_size=-1, th=-1e30),
                   dict(scale=1.3, min_size=30, max_size=60, th=0.5)):
            _same(c.detect(img, **kw), oracle.detect(ho, img, **kw))
        tn, ts, lv = c.trace(img, leaf_range=(0, 500))
        on, os_, olv = oracle.trace(ho, img, leaf_range=(0, 500))
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        np.testing.assert_array_equal(lv, olv)
    c.close(); oracle.release(ho)


@pytest.mark.parametrize("stage_kernels", [False, True], ids=["k3_cascade", "stage_kernels"])
@pytest.mark.parametrize("n_frames", [9, 130])
def test_scan_plus_planes_batches(oracle, tmp_path, n_frames, stage_kernels):
    """Stage 0 from the LUT scan, stages >= 1 sampling the h / q planes, in batch mode (cohort-staged stage 0; 130
    frames: the size at which a scale-0 model would be copied in chunks -- a model with planes must not be).
    c/jda.c:340-354, 385-394."""
    path = synth.write_model(str(tmp_path / "syn.model"), seed=6, mode="reject", scales=(0, 1, 2), coord_max=0.45,
                             scales_by_stage={0: (0,)})
    # (small frames: such a batch takes k3_cascade for the stages >= 1 unless the stage kernels -- here their h / q
    # instantiation, k3_walk<., true> -- are forced)
    os.environ["JDA_B200_STAGE_MIN_WINDOWS"] = "0" if stage_kernels else "1000000000000"
    os.environ["JDA_B200_FORCE_PLAN"] = "throughput"   # (9 small frames would otherwise take the latency plan)
    try:
        c = api.Cascador(path, double=True)
    finally:
        del os.environ["JDA_B200_STAGE_MIN_WINDOWS"], os.environ["JDA_B200_FORCE_PLAN"]
    ho = oracle.load(path, True)
    frames = synth.make_frames("facemix", n_frames, 112, 90, seed0=300)
    res = c.detect_batch(frames, th=-1e30, flags=api.RAW_HITS)
    st = c.last_stats
    assert st["scan_launches"] >= 1 and st["resize_launches"] == 1, st
    assert (st["cascade_launches"] > 2) == stage_kernels, st
    hits = 0
    for f in range(n_frames):
        ob, osc, osh, ost = oracle.detect_raw(ho, frames[f], th=-1e30)
        assert ost["ub_reads"] == 0
        _same(res[f], (ob, osc, osh))
        hits += len(osc)
    assert hits > 0
    # mixed sizes with such a model go shape group by shape group (the canvas path needs a plane-free model)
    sizes = [(112, 90), (90, 112), (112, 90), (64, 48), (90, 112)]
    imgs = [synth.facemix_frame(400 + i, w, h) for i, (w, h) in enumerate(sizes)]
    got = c.detect_mixed(imgs, th=-1e30)
    for g, img in zip(got, imgs):
        _same(g, oracle.detect(ho, img, th=-1e30))
    c.close(); oracle.release(ho)


def test_more_than_twenty_pyramid_levels(casc, oracle, oracle_shipped):
    """1080p at scale 1.2 visits 22 window sizes, 4K at the reference's own 1.25 visits 21 (c/jda.c:331-332):
    round 1 refused both (kMaxLevels = 20)."""
    assert len(api.levels(1920, 1080, 1.2, 24, -1)) == 22 and len(api.levels(3840, 2160, 1.25, 24, -1)) == 21
    img = synth.facemix_frame(77, 1920, 1080)
    _same(casc.detect(img, scale=1.2), oracle.detect(oracle_shipped, img, scale=1.2))
    # 4K: flat background (cheap for the CPU oracle) with faces pasted at several sizes
    big = np.full((2160, 3840), 100, np.uint8)
    small = synth.face_canvas()
    big[200:200 + 480, 300:300 + 640] = small
    big[1200:1200 + 960, 2000:2000 + 1280] = np.kron(small, np.ones((2, 2), np.uint8))
    got = casc.detect(big)
    _same(got, oracle.detect(oracle_shipped, big))
    assert len(got[1]) >= 2


DIMS = [dict(T=2, K=33, L=5), dict(T=3, K=96, L=40), dict(T=1, K=540, L=27), dict(T=6, K=64, L=16),
        dict(T=2, K=1000, L=8)]        # K = 1000: the stage-0 table does not fit shared memory -> generic kernel only


@pytest.mark.parametrize("dims", DIMS, ids=lambda d: "T%d_K%d_L%d" % (d["T"], d["K"], d["L"]))
@pytest.mark.parametrize("mode", ["reject", "passall"])
def test_model_dimensions_from_the_header(oracle, tmp_path, dims, mode):
    """T, K, landmark_n come from the model header, not from compile-time macros
    (c/jda.c:24-32 fixes them at 5 / 540 / 27).  Oracle = the restatement, which reads the header too."""
    if mode == "passall" and dims["K"] * dims["T"] > 1200:
        pytest.skip("pass-all on a deep model is covered by the default dimensions")
    path = synth.write_model(str(tmp_path / "dims.model"), seed=7 + dims["K"], mode=mode, norm_every=270 if dims["K"] > 900 else 10, **dims)
    c = api.Cascador(path, double=True)
    ho = oracle.load(path, True)
    assert (c.T, c.K, c.L) == (dims["T"], dims["K"], dims["L"]) == oracle.dims(ho)[:3]
    frames = [synth.blur_frame(9, 96, 80), synth.noise_frame(5, 70, 61), synth.facemix_frame(3, 200, 150)]
    for img in frames:
        for kw in (dict(th=-1e30), dict(scale=1.3, min_size=30, max_size=60, th=0.0)):
            _same(c.detect(img, **kw), oracle.detect(ho, img, **kw))
    batch = synth.make_frames("facemix", 9, 120, 90, seed0=5)      # > 4 frames: throughput plan + staged stage 0
    for g, f in zip(c.detect_batch(batch, th=-1e30, flags=api.RAW_HITS), batch):
        ob, osc, osh, _ = oracle.detect_raw(ho, f, th=-1e30)
        _same(g, (ob, osc, osh))
    tn, ts, lv = c.trace(frames[0], leaf_range=(0, 300))
    on, os_, olv = oracle.trace(ho, frames[0], leaf_range=(0, 300))
    np.testing.assert_array_equal(tn, on)
    np.testing.assert_array_equal(_bits(ts), _bits(os_))
    np.testing.assert_array_equal(lv, olv)
    out = tmp_path / "rt.model"
    c.save_f32(str(out)); oracle.save_f32(ho, str(tmp_path / "rt_o.model"))
    assert out.read_bytes() == (tmp_path / "rt_o.model").read_bytes()
    c.close(); oracle.release(ho)


@pytest.mark.parametrize("depth", [2, 3, 5, 6])
def test_tree_depth_from_the_header(oracle, tmp_path, depth):
    """tree_depth 2..6 (c/jda.c:24-32 compiles in 4): every stage through the generic cascade
    kernel with run-time node / leaf counts.  Oracle = the restatement (the reference binary only loads depth 4)."""
    path = synth.write_model(str(tmp_path / "d.model"), seed=50 + depth, T=3, K=70, L=11, depth=depth, norm_every=9)
    c = api.Cascador(path, double=True)
    ho = oracle.load(path, True)
    assert c.depth == depth
    frames = [synth.blur_frame(9, 96, 80), synth.facemix_frame(3, 200, 150)]
    for img in frames:
        for kw in (dict(th=-1e30), dict(scale=1.3, min_size=30, max_size=60, th=0.0)):
            _same(c.detect(img, **kw), oracle.detect(ho, img, **kw))
    batch = synth.make_frames("facemix", 7, 120, 90, seed0=5)
    for g, f in zip(c.detect_batch(batch, th=-1e30, flags=api.RAW_HITS, t_limit=1, k_limit=33), batch):
        ob, osc, osh, _ = oracle.detect_raw(ho, f, th=-1e30, t_limit=1, k_limit=33)
        _same(g, (ob, osc, osh))
    tn, ts, lv = c.trace(frames[0], leaf_range=(0, 400))
    on, os_, olv = oracle.trace(ho, frames[0], leaf_range=(0, 400))
    np.testing.assert_array_equal(tn, on)
    np.testing.assert_array_equal(_bits(ts), _bits(os_))
    np.testing.assert_array_equal(lv, olv)
    c.close(); oracle.release(ho)


def test_handles_with_different_models_do_not_disturb_each_other(casc, oracle, oracle_shipped, tmp_path):
    """Kernel attributes (dynamic shared-memory limits) belong to the device, not to a handle: a handle with a small
    model initialised between two calls of a handle with a large one must not lower the limit under it (r2 bug)."""
    img = synth.face_canvas()
    want = oracle.detect(oracle_shipped, img)
    _same(casc.detect(img), want)
    small = api.Cascador(synth.write_model(str(tmp_path / "small.model"), seed=3, T=2, K=40, L=6), double=True)
    small.detect(synth.noise_frame(1, 64, 48))          # first use: this handle's ctx_init sets the attributes again
    _same(casc.detect(img), want)
    _same(casc.detect_batch(np.stack([img] * 6))[3], want)
    small.close()


def test_against_reference_library_directly(reflib, tmp_path):
    path = synth.write_model(str(tmp_path / "syn.model"), seed=21, mode="reject")
    c = api.Cascador(path, double=True)
    hr = reflib.load(path, True)
    for img in (synth.blur_frame(31, 200, 160), synth.facemix_frame(32, 333, 250)):
        _same(c.detect(img, th=-1e30), reflib.detect(hr, img, th=-1e30))
    c.close(); reflib.release(hr)


def test_resize_planes_byte_identical(casc, oracle):
    img = synth.facemix_frame(2)
    r = np.float32(1.0) / np.sqrt(np.float32(2.0))
    for dw, dh in [(int(np.float32(640) * r), int(np.float32(480) * r)), (320, 240), (101, 77)]:
        np.testing.assert_array_equal(casc.resize(img, dw, dh), oracle.resize(img, dw, dh))


# ---- batch / device-resident / mining entry points --------------------------------------------------

def test_batch_equals_per_frame(casc, oracle, oracle_shipped):
    frames = synth.make_frames("mix", 9, seed0=40)
    res = casc.detect_batch(frames, max_size=192, th=-0.5)
    assert casc.last_stats["windows"] == 9 * 169236
    for f in range(9):
        _same(res[f], oracle.detect(oracle_shipped, frames[f], max_size=192, th=-0.5))


def test_flat_batch_result_equals_per_frame_results(casc):
    """jdaB200DetectBatchFlat: one result for the batch = the per-frame jdaResults back to back"""
    from jda_b200 import shard
    frames = synth.make_frames("facemix", 24, 320, 240, seed0=60)
    for kw in (dict(th=-0.5), dict(th=0.0, flags=api.RAW_HITS | api.NO_FINAL_TH, t_limit=2)):
        per = casc.detect_batch(frames, **kw)
        counts, boxes, scores, shapes = casc.detect_batch(frames, flat=True, **kw)
        assert counts.tolist() == [len(r[1]) for r in per] and counts.sum() == len(scores) >= 3
        _same((boxes, scores, shapes), (np.concatenate([r[0] for r in per]), np.concatenate([r[1] for r in per]),
                                        np.concatenate([r[2] for r in per])))
        a = shard.pack_records(per, frame0=100, landmark_n=casc.L)
        b = shard.pack_records_flat(counts, boxes, scores, shapes, frame0=100)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    empty = casc.detect_batch(synth.make_frames("noise", 3, 64, 48), flat=True)
    assert empty[0].tolist() == [0, 0, 0] and empty[1].shape == (0, 3) and empty[3].shape == (0, 54)


def test_device_resident_frames(casc):
    import torch
    frames = synth.make_frames("facemix", 6, seed0=60)
    host = casc.detect_batch(frames, max_size=192, th=-0.5)
    d = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    dev = casc.detect_batch(None, device_ptr=d.data_ptr(), shape=tuple(d.shape), max_size=192, th=-0.5)
    for a, b in zip(host, dev):
        _same(a, b)
    # odd pitch: TMA not applicable, plain-load path must give the same answer
    fr = synth.make_frames("facemix", 3, 451, 333, seed0=70)
    host = casc.detect_batch(fr, th=-0.5)
    d = torch.from_numpy(fr).cuda()
    torch.cuda.synchronize()
    dev = casc.detect_batch(None, device_ptr=d.data_ptr(), shape=tuple(d.shape), th=-0.5)
    for a, b in zip(host, dev):
        _same(a, b)


def test_mixed_size_frames_fddb_shaped(casc, oracle, oracle_shipped):
    """BASELINE config 4 shape: FDDB-like frame sizes through jdaB200DetectMixed -- every frame in its slot of a
    common canvas, one launch per kernel, each frame keeps the windows of its own size (c/jda.c:320-339)"""
    frames = [synth.facemix_frame(300 + i, *synth.fddb_shape(i % 5)) for i in range(12)]
    want = [oracle.detect(oracle_shipped, f, th=-0.5) for f in frames]
    got = casc.detect_many(frames, th=-0.5)
    assert casc.last_stats["windows"] == sum(api.count_windows(f.shape[1], f.shape[0]) for f in frames)
    assert casc.last_stats["scan_launches"] == 1
    assert sum(len(w[1]) for w in want) >= 3
    for g, w in zip(got, want):
        _same(g, w)
    grouped = casc.detect_many(frames, group=True, th=-0.5)   # the per-shape batches give the same answer
    assert casc.last_stats["scan_launches"] == len({f.shape for f in frames})
    for g, w in zip(grouped, want):
        _same(g, w)


def test_mixed_sizes_edge_cases(casc, oracle, oracle_shipped):
    """canvas corner cases: portrait + landscape in one call (canvas larger than either), frames too small for any
    window, a strided view as input, an explicit max_size, tiny and empty batches, the latency plan (<= 4 frames)"""
    big = synth.facemix_frame(41, 300, 200)
    wide = np.zeros((120, 400), np.uint8); wide[:] = synth.facemix_frame(42, 400, 120)
    tall = synth.facemix_frame(43, 130, 380)
    tiny = synth.noise_frame(3, 23, 40)           # < 24 px wide: no windows (c/jda.c:320-322)
    exact = synth.blur_frame(5, 24, 24)           # exactly one window
    view = synth.facemix_frame(44, 512, 256)[16:200, 40:300]   # non-contiguous rows (pitch 512, width 260)
    frames = [big, wide, tall, tiny, exact, view]
    for kw in (dict(th=-0.5), dict(th=-0.5, max_size=100), dict(scale=1.2, min_size=40, th=-1.0)):
        got = casc.detect_mixed(frames, **kw)
        st = casc.last_stats
        assert st["windows"] == sum(api.count_windows(f.shape[1], f.shape[0], kw.get("scale", 1.25),
                                                       kw.get("min_size", 24), kw.get("max_size", -1)) for f in frames)
        for f, g in zip(frames, got):
            _same(g, oracle.detect(oracle_shipped, np.ascontiguousarray(f), **kw))
    for g, f in zip(casc.detect_mixed(frames, th=-0.5, flags=api.NO_TMA), frames):     # tiles filled by plain loads
        _same(g, oracle.detect(oracle_shipped, np.ascontiguousarray(f), th=-0.5))
    assert casc.detect_mixed([]) == []
    got = casc.detect_mixed([tiny])
    assert len(got) == 1 and len(got[0][1]) == 0
    two = casc.detect_mixed([tall, big], th=-0.5)              # <= 4 frames: latency tile plan
    _same(two[0], oracle.detect(oracle_shipped, tall, th=-0.5))
    _same(two[1], oracle.detect(oracle_shipped, big, th=-0.5))
    # mining flags ride along: first stage only, every survivor, window-normalised shapes
    raw = casc.detect_mixed([wide, tall], t_limit=1, flags=api.RAW_HITS | api.NO_FINAL_TH)
    for f, g in zip([wide, tall], raw):
        ob, osc, osh, _ = oracle.detect_raw(oracle_shipped, f, t_limit=1, use_th=False)
        _same(g, (ob, osc, osh))


def test_mixed_sizes_large_batch_chunked(casc, oracle, oracle_shipped):
    """>= 128 mixed frames: the per-frame copies are split in chunks that overlap the scans; sampled frames vs oracle"""
    shapes = [synth.fddb_shape(s) for s in range(7)]
    frames = [synth.facemix_frame(800 + i, *[d // 2 for d in shapes[i % 7]]) for i in range(140)]
    got = casc.detect_mixed(frames, th=-0.5)
    assert casc.last_stats["scan_launches"] == 4
    assert casc.last_stats["windows"] == sum(api.count_windows(f.shape[1], f.shape[0]) for f in frames)
    for i in (0, 1, 2, 3, 4, 5, 6, 69, 139):
        _same(got[i], oracle.detect(oracle_shipped, frames[i], th=-0.5))
    again = casc.detect_mixed(frames[::-1], th=-0.5)
    for a, b in zip(got, again[::-1]):
        _same(a, b)


def test_mixed_sizes_without_the_canvas_path(casc, oracle, oracle_shipped, tmp_path):
    """models / flags that cannot share a canvas launch (h / q planes differ per frame size, or no stage-0 scan):
    jdaB200DetectMixed then runs one batch per distinct shape internally -- same API, same answers"""
    frames = [synth.blur_frame(9, 96, 80), synth.noise_frame(5, 70, 61), synth.blur_frame(10, 96, 80),
              synth.facemix_frame(4, 120, 90), synth.noise_frame(2, 20, 30)]
    path = synth.write_model(str(tmp_path / "scaled.model"), seed=3, mode="reject", scales=(0, 1, 2), coord_max=0.45)
    c = api.Cascador(path, double=True)
    ho = oracle.load(path, True)
    got = c.detect_mixed(frames, th=-1e30)
    assert c.last_stats["scan_launches"] == 0 and c.last_stats["cascade_launches"] >= 3
    assert c.last_stats["windows"] == sum(api.count_windows(f.shape[1], f.shape[0]) for f in frames)
    for g, f in zip(got, frames):
        _same(g, oracle.detect(ho, f, th=-1e30))
    c.close(); oracle.release(ho)
    big = [synth.facemix_frame(41, 300, 200), synth.facemix_frame(43, 130, 380), synth.facemix_frame(41, 300, 200)]
    got = casc.detect_mixed(big, th=-0.5, flags=api.NO_STAGE0_SCAN)
    assert casc.last_stats["scan_launches"] == 0
    for g, f in zip(got, big):
        _same(g, oracle.detect(oracle_shipped, f, th=-0.5))


def test_chunked_host_batch_equals_resident(casc, oracle, oracle_shipped):
    """>= 128 host frames are copied and scanned in overlapping chunks; same answer as one resident batch"""
    import torch
    frames = synth.make_frames("facemix", 131, 200, 150, seed0=900)
    frames[7, 20:128, 30:141] = np.load(os.path.join(os.path.dirname(__file__), "golden", "face_111x108.npy"))
    host = casc.detect_batch(frames, th=-0.5)
    assert casc.last_stats["scan_launches"] == 4
    d = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    dev = casc.detect_batch(None, device_ptr=d.data_ptr(), shape=tuple(d.shape), th=-0.5)
    assert casc.last_stats["scan_launches"] == 1
    assert sum(len(r[1]) for r in host) >= 1
    for a, b in zip(host, dev):
        _same(a, b)
    for f in (0, 7, 65, 130):
        _same(host[f], oracle.detect(oracle_shipped, frames[f], th=-0.5))


@pytest.mark.parametrize("t_limit", [1, 2, 5])
def test_mining_mode_truncated_cascade(casc, oracle, oracle_shipped, t_limit):
    """Validate()'s partial cascade (src/jda/cascador.cpp:178-197): first t stages, every survivor
    emitted with its score and window-normalised shape, no NMS."""
    img = synth.face_canvas()
    got = casc.detect_batch(img[None], t_limit=t_limit, flags=api.RAW_HITS | api.NO_FINAL_TH)[0]
    ob, osc, osh, st = oracle.detect_raw(oracle_shipped, img, t_limit=t_limit, use_th=False)
    assert len(osc) == st["stage_survivors"][t_limit - 1] > 0
    _same(got, (ob, osc, osh))


@pytest.mark.parametrize("tk", [(0, 18), (0, 540), (2, 101), (1, 1), (4, 270)], ids=lambda tk: "t%d_k%d" % tk)
@pytest.mark.parametrize("n_frames", [1, 6])
def test_mining_mode_cart_granular(casc, casc_stages, oracle, oracle_shipped, tk, n_frames):
    """Validate() while a stage is being trained (src/jda/cascador.cpp:178-209, caller btcart.cpp:146-152):
    t full stages, then carts [0, k) of stage t, no regression after them.  1 frame: latency plan (k3_cascade
    redoes stage 0); 6 frames: throughput plan, once with k3_cascade for the stages >= 1 (what a batch this small takes)
    and once with the stage kernels (k2_scan -> k3_regress -> k3_walk ... -> k3_emit)."""
    t, k = tk
    frames = np.stack([synth.face_canvas()] + [synth.facemix_frame(90 + i) for i in range(n_frames - 1)])
    got = casc.detect_batch(frames, t_limit=t, k_limit=k, flags=api.RAW_HITS | api.NO_FINAL_TH)
    if n_frames > 1:
        forced = casc_stages.detect_batch(frames, t_limit=t, k_limit=k, flags=api.RAW_HITS | api.NO_FINAL_TH)
        for a, b in zip(got, forced):
            _same(a, b)
    total = 0
    for f in range(n_frames):
        ob, osc, osh, st = oracle.detect_raw(oracle_shipped, frames[f], t_limit=t, k_limit=k, use_th=False)
        _same(got[f], (ob, osc, osh))
        total += len(osc)
    assert total > 0
    if n_frames == 1:  # per-window trace: carts evaluated stop at t*K + k, exit score, leaves of the partial stage
        nwin = api.count_windows(640, 480)
        rng = (nwin - 2500, nwin)
        tn, ts, lv = casc.trace(frames[0], t_limit=t, k_limit=k, leaf_range=rng)
        on, os_, olv = oracle.trace(oracle_shipped, frames[0], t_limit=t, k_limit=k, leaf_range=rng)
        assert on.max() == t * 540 + k
        np.testing.assert_array_equal(tn, on)
        np.testing.assert_array_equal(_bits(ts), _bits(os_))
        np.testing.assert_array_equal(lv, olv)


def test_submit_collect_pipeline(casc):
    """jdaB200Submit / jdaB200Collect: two batches in flight give the results of the synchronous call, whatever the
    order of sizes and pyramids; overflowing queues are re-run at collect time; misuse is refused, not undefined."""
    batches = [synth.make_frames("mix", 9, seed0=800), synth.make_frames("facemix", 7, 320, 240, seed0=810),
               synth.make_frames("mix", 9, seed0=820), synth.make_frames("blur6", 3, seed0=830),
               synth.make_frames("facemix", 140, 200, 150, seed0=840)]
    kws = [dict(max_size=192, th=-0.5), dict(th=0.0), dict(max_size=192, th=-0.5), dict(scale=1.3, min_size=30, th=-1.0),
           dict(th=-0.5)]
    want = [casc.detect_batch(b, flat=True, **kw) for b, kw in zip(batches, kws)]
    c = api.Cascador(SHIPPED_F32, double=False)
    tickets = [c.submit(batches[0], **kws[0])]
    got = []
    for i in range(1, len(batches)):
        tickets.append(c.submit(batches[i], **kws[i]))      # batch i goes in while batch i - 1 is still running
        got.append(c.collect(tickets[i - 1]))
        assert c.last_stats["windows"] > 0 and c.last_stats["scan_launches"] >= 1
    got.append(c.collect(tickets[-1]))
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            np.testing.assert_array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
    # misuse
    t0 = c.submit(batches[0], **kws[0])
    t1 = c.submit(batches[2], **kws[2])
    with pytest.raises(RuntimeError, match="already in flight"):
        c.submit(batches[0], **kws[0])
    with pytest.raises(RuntimeError, match="waiting for jdaB200Collect"):
        c.detect(batches[0][0])
    c.collect(t0); c.collect(t1)
    with pytest.raises(RuntimeError, match="not in flight"):
        c.collect(t1)
    _same(c.detect(batches[0][0], max_size=192, th=-0.5), casc.detect(batches[0][0], max_size=192, th=-0.5))
    c.close()
    # queues that overflow: the batch is re-run when it is collected
    os.environ["JDA_B200_TINY_QUEUES"] = "1"
    try:
        c2 = api.Cascador(SHIPPED_F32, double=False)
    finally:
        del os.environ["JDA_B200_TINY_QUEUES"]
    ta = c2.submit(batches[0], **kws[0])
    tb = c2.submit(batches[2], **kws[2])
    for g, w in ((c2.collect(ta), want[0]), (c2.collect(tb), want[2])):
        for a, b in zip(g, w):
            np.testing.assert_array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))
    c2.close()


def test_concurrent_callers_and_changing_arguments(casc, oracle, oracle_shipped):
    """jdaDetect is re-entrant in the reference: threads share one handle; calls with different
    sizes / pyramids interleave (the per-handle geometry cache is rebuilt as needed)."""
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(synth.face_canvas(), dict(th=0.0)), (synth.facemix_frame(5), dict(scale=1.2, min_size=30, max_size=300, th=-1.0)),
            (synth.facemix_frame(7, *synth.fddb_shape(7)), dict(th=0.0)), (synth.blur_frame(2, 30, 27), dict(th=-5.0)),
            (synth.facemix_frame(3), dict(max_size=192, th=0.0))] * 3
    want = [oracle.detect(oracle_shipped, img, **kw) for img, kw in jobs[:5]] * 3
    with ThreadPoolExecutor(6) as ex:
        got = list(ex.map(lambda j: casc.detect(j[0], **j[1]), jobs))
    for g, w in zip(got, want):
        _same(g, w)
    c2 = api.Cascador(SHIPPED_F32, double=False)       # a second handle on the same device
    with ThreadPoolExecutor(4) as ex:
        got = list(ex.map(lambda a: (casc if a[0] % 2 else c2).detect(a[1][0], **a[1][1]), enumerate(jobs)))
    for g, w in zip(got, want):
        _same(g, w)
    c2.close()


# ---- edges ------------------------------------------------------------------------------------

def test_edge_cases(casc, oracle, oracle_shipped):
    tiny = synth.blur_frame(2, 30, 27)
    for img, kw in [(tiny, dict(th=-5.0)),
                    (synth.blur_frame(3, 24, 24), dict(th=-1e30)),      # exactly one window
                    (synth.blur_frame(3, 23, 64), dict()),               # too small: no windows
                    (synth.noise_frame(1, 100, 100), dict(scale=1.0)),   # reference would never return
                    (synth.noise_frame(1, 100, 100), dict(min_size=90, max_size=50)),
                    (synth.face_canvas(), dict(scale=2.0, th=-1.0)),
                    (synth.face_canvas(), dict(scale=1.07, min_size=150, th=-1.0))]:
        _same(casc.detect(img, **kw), oracle.detect(oracle_shipped, img, **kw))
    assert casc.detect(synth.blur_frame(3, 23, 64))[0].shape == (0, 3)


# ---- full-size, size-independent properties -------------------------------------------------------

def test_full_size_batch_properties(casc, oracle, oracle_shipped):
    """BASELINE config 2 shape (VGA, 3-octave) on a 96-frame slice: window count is exact, the run is
    idempotent, batch order does not matter, and sampled frames equal the oracle."""
    frames = synth.make_frames("mix", 96, seed0=1000)
    a = casc.detect_batch(frames, max_size=192, th=0.0)
    st = dict(casc.last_stats)
    assert st["windows"] == 96 * 169236 and st["n_levels"] == 10
    b = casc.detect_batch(frames, max_size=192, th=0.0)
    for x, y in zip(a, b):
        _same(x, y)
    perm = np.random.default_rng(0).permutation(96)
    c = casc.detect_batch(frames[perm], max_size=192, th=0.0)
    for i, p in enumerate(perm):
        _same(c[i], a[p])
    s0 = 0
    # every one of the 96 frames against the oracle (the restatement releases the GIL inside ctypes: host threads)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        want = list(ex.map(lambda f: oracle.detect(oracle_shipped, frames[f], max_size=192, th=0.0), range(96)))
    for f in range(96):
        _same(a[f], want[f])
    assert sum(len(w[1]) for w in want) > 20
    raw = casc.detect_batch(frames[:4], max_size=192, flags=api.RAW_HITS | api.NO_FINAL_TH)
    for f in range(4):
        ob, osc, osh, ost = oracle.detect_raw(oracle_shipped, frames[f], max_size=192, use_th=False)
        _same(raw[f], (ob, osc, osh))
        s0 += ost["stage_survivors"][0]
    assert casc.last_stats["stage0_survivors"] == s0


def test_1080p_five_octave_frames(casc, oracle, oracle_shipped):
    """BASELINE config 3 (1920x1080, min 24, max 768: 16 levels, 1,245,202 windows per frame) through plain jdaDetect:
    four frames of three kinds against the oracle, beside the golden reference output of `hd_blur`"""
    from concurrent.futures import ThreadPoolExecutor
    big = np.full((1080, 1920), 100, np.uint8)                         # flat background, the known-answer canvas at 2x
    big[60:60 + 960, 300:300 + 1280] = np.kron(synth.face_canvas(), np.ones((2, 2), np.uint8))
    frames = [synth.facemix_frame(7000, 1920, 1080), synth.facemix_frame(7001, 1920, 1080), synth.blur_frame(7002, 1920, 1080), big]
    kw = dict(scale=1.25, min_size=24, max_size=768, th=0.0)
    assert api.count_windows(1920, 1080, 1.25, 24, 768) == 1245202
    with ThreadPoolExecutor(4) as ex:
        want = list(ex.map(lambda f: oracle.detect(oracle_shipped, f, **kw), frames))
    for f, w in zip(frames, want):
        _same(casc.detect(f, **kw), w)
    assert sum(len(w[1]) for w in want) >= 3
    got = casc.detect_batch(np.stack(frames + frames[:2]), **kw)          # and as a batch (throughput tile plan)
    for g, w in zip(got, want + want[:2]):
        _same(g, w)


def test_cli_detect_fddb_format(tmp_path, oracle, oracle_shipped):
    from jda_b200.__main__ import main
    img = synth.face_canvas()
    np.save(tmp_path / "a.npy", img)
    with open(tmp_path / "b.pgm", "wb") as f:
        f.write(b"P5\n640 480\n255\n" + img.tobytes())
    out = tmp_path / "fold-01-out.txt"
    assert main(["detect", SHIPPED_F32, str(tmp_path / "a.npy"), str(tmp_path / "b.pgm"), "--float", "--fddb-out", str(out)]) == 0
    lines = out.read_text().split("\n")
    b, s, _ = oracle.detect(oracle_shipped, img)
    assert lines[0] == str(tmp_path / "a") and lines[1] == "2"
    assert lines[2] == "%d %d %d %d %f" % (b[0][0], b[0][1], b[0][2], b[0][2], s[0])
    assert lines[4] == str(tmp_path / "b") and lines[5] == "2"


def test_queue_overflow_grows_and_retries(oracle, oracle_shipped):
    """survivor / hit queues that are too small are detected from the device counters, grown and the batch re-run"""
    os.environ["JDA_B200_TINY_QUEUES"] = "1"
    try:
        c = api.Cascador(SHIPPED_F32, double=False)
        frames = np.stack([synth.face_canvas(), synth.facemix_frame(5), synth.noise_frame(2)])
        got = c.detect_batch(frames, th=-1.0)
        assert c.last_stats["scan_launches"] >= 2 and c.last_stats["stage0_survivors"] > 8
        for f in range(3):
            _same(got[f], oracle.detect(oracle_shipped, frames[f], th=-1.0))
        raw = c.detect_batch(frames[:1], flags=api.RAW_HITS | api.NO_FINAL_TH)[0]
        ob, osc, osh, _ = oracle.detect_raw(oracle_shipped, frames[0], use_th=False)
        _same(raw, (ob, osc, osh))
        c.close()
    finally:
        del os.environ["JDA_B200_TINY_QUEUES"]


# ---- stages >= 1: the stage-synchronous kernels (k3_walk / k3_regress / k3_emit) against the one-warp-per-window
# ---- kernel (k3_cascade) and the round-1 regression kernel (k3_stage0)

@pytest.mark.parametrize("mode", ["full", "t2_k101", "t3"])
def test_stage_kernels_equal_one_warp_per_window(casc_stages, oracle, oracle_shipped, mode):
    """A 40-frame batch (throughput plan) with many deep survivors: faces pass all five stages, blurred frames die in
    stages 1-3.  Stage kernels (forced: a batch this small takes k3_cascade by default) = k3_walk + k3_regress + k3_emit; JDA_B200_NO_STAGE_KERNELS = k3_cascade for stages >= 1;
    JDA_B200_OLD_REGRESS = k3_stage0 for every regression.  All three give the same bits, and frames 0 / 1 / 21 equal
    the oracle's."""
    frames = np.stack([synth.face_canvas()] + [synth.facemix_frame(300 + i) for i in range(19)] +
                      list(synth.make_frames("blur6", 10, seed0=340)) + list(synth.make_frames("mix", 10, seed0=350)))
    kw = dict(flags=api.RAW_HITS | api.NO_FINAL_TH)
    okw = dict(use_th=False)
    if mode == "t2_k101":
        kw.update(t_limit=2, k_limit=101); okw.update(t_limit=2, k_limit=101)
    elif mode == "t3":
        kw.update(t_limit=3); okw.update(t_limit=3)
    got = casc_stages.detect_batch(frames, **kw)
    launches = casc_stages.last_stats["cascade_launches"]
    c1 = _with_env("JDA_B200_NO_STAGE_KERNELS")
    c2 = _with_env("JDA_B200_OLD_REGRESS", JDA_B200_STAGE_MIN_WINDOWS=0)
    try:
        old = c1.detect_batch(frames, **kw)
        assert c1.last_stats["cascade_launches"] == 2 < launches   # (k3_stage0 + k3_cascade) against one pair per stage + emit
        oldr = c2.detect_batch(frames, **kw)
    finally:
        c1.close(); c2.close()
    assert sum(len(g[1]) for g in got) > 50
    for a, b, d in zip(got, old, oldr):
        _same(a, b)
        _same(a, d)
    for f in (0, 1, 21):
        ob, osc, osh, _ = oracle.detect_raw(oracle_shipped, frames[f], **okw)
        _same(got[f], (ob, osc, osh))


def test_concurrent_jdadetect_calls_are_coalesced(oracle, oracle_shipped):
    """Sixteen host threads on one handle, every one calling the reference's own entry point (jdaDetect) on frames of
    different sizes: calls that arrive while another is running are served together as one mixed-size batch -- and every
    caller still gets, bit for bit, what its frame gives alone (compared with the serial results of the same handle and,
    for a sample, with the oracle).  Calls with other parameters are never merged into the same batch."""
    from concurrent.futures import ThreadPoolExecutor
    c = api.Cascador(SHIPPED_F32, double=False)
    frames = [synth.facemix_frame(500 + i, *synth.fddb_shape(i)) for i in range(40)] + \
             [synth.face_canvas(), synth.blur_frame(2, 30, 27), synth.noise_frame(3, 23, 64)] + \
             [synth.facemix_frame(560 + i) for i in range(21)]
    kws = [dict(th=-0.5) if i % 7 else dict(scale=1.3, min_size=30, th=-1.0) for i in range(len(frames))]
    want = [c.detect(f, **kw) for f, kw in zip(frames, kws)]
    calls0, batches0, _ = c.coalescing_stats()
    assert calls0 == batches0 == len(frames)            # serial calls: one batch each
    with ThreadPoolExecutor(16) as ex:
        got = list(ex.map(lambda a: c.detect(a[0], **a[1]), zip(frames, kws)))
    for g, w in zip(got, want):
        _same(g, w)
    calls, batches, largest = c.coalescing_stats()
    assert calls - calls0 == len(frames)
    assert batches - batches0 < len(frames) and largest >= 2, (calls, batches, largest)
    for i in (0, 7, 40, 41, 42, 50):
        _same(got[i], oracle.detect(oracle_shipped, frames[i], **kws[i]))
    c.close()
