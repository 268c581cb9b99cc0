"""CPU-side checks of libjda_b200.so: it loads, exports the whole C ABI, and its host logic
(model files, level enumeration, NMS) agrees with the oracle.  No compute calls need a GPU."""
import os
import re

import numpy as np
import pytest

from jda_b200 import api, synth
from tests.conftest import ROOT, SHIPPED_F32


def test_library_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "jda_b200.h")).read()
    declared = re.findall(r"JDA_API\s+[\w\s\*]+?\b(jda\w+)\s*\(", hdr)
    assert len(declared) >= 27 and set(declared) == set(api.EXPORTS)
    L = ctypes.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    # the six symbols of the reference's c/jda.h are all there under their real names
    for name in ("jdaCascadorCreateDouble", "jdaCascadorCreateFloat", "jdaCascadorSerializeTo",
                 "jdaCascadorRelease", "jdaDetect", "jdaResultRelease"):
        assert name in declared


def test_result_struct_layout_matches_reference_header():
    import ctypes
    # c/jda.h:18-24: two ints + three pointers = 32 bytes on LP64
    assert ctypes.sizeof(api._Result) == 32
    assert api._Result.bboxes.offset == 8 and api._Result.shapes.offset == 16 and api._Result.scores.offset == 24


def test_model_roundtrip_bytes(tmp_path):
    c = api.Cascador(SHIPPED_F32, double=False)
    assert (c.T, c.K, c.L, c.depth) == (5, 540, 27, 4)
    out = tmp_path / "rt.model"
    c.save_f32(str(out))
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()
    # double flavour -> same in-memory model -> same float file
    wide = synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model"))
    c2 = api.Cascador(wide, double=True)
    out2 = tmp_path / "rt2.model"
    c2.save_f32(str(out2))
    assert out2.read_bytes() == out.read_bytes()
    c.close(); c2.close()


def test_synthetic_model_serialiser_matches_oracle(oracle, tmp_path):
    path = synth.write_model(str(tmp_path / "s.model"), seed=11, scales=(0, 1, 2), coord_max=0.45)
    c = api.Cascador(path, double=True)
    h = oracle.load(path, True)
    a, b = tmp_path / "a", tmp_path / "b"
    c.save_f32(str(a)); oracle.save_f32(h, str(b))
    assert a.read_bytes() == b.read_bytes()
    c.close(); oracle.release(h)


def test_bad_models_are_rejected(tmp_path):
    with pytest.raises(RuntimeError):
        api.Cascador(str(tmp_path / "missing.model"))
    p = tmp_path / "short.model"
    p.write_bytes(open(SHIPPED_F32, "rb").read()[:100000])
    with pytest.raises(RuntimeError):
        api.Cascador(str(p), double=False)
    q = tmp_path / "depth5.model"          # depth 5 is a legal header now, but this file holds depth-4 carts: short read
    b = bytearray(open(SHIPPED_F32, "rb").read())
    b[16:20] = (5).to_bytes(4, "little")
    q.write_bytes(bytes(b))
    with pytest.raises(RuntimeError):
        api.Cascador(str(q), double=False)
    b[16:20] = (7).to_bytes(4, "little")      # outside 2..6
    q.write_bytes(bytes(b))
    with pytest.raises(RuntimeError):
        api.Cascador(str(q), double=False)


@pytest.mark.parametrize("depth", [2, 3, 5, 6])
def test_tree_depth_comes_from_the_header(oracle, tmp_path, depth):
    """c/jda.c:24-32 fixes JDA_TREE_DEPTH = 4 at compile time; here the loader takes it from the
    header (README.md:84-111) and the serialiser writes it back.  Oracle = the restatement, which does the same."""
    path = synth.write_model(str(tmp_path / "d.model"), seed=40 + depth, T=2, K=50, L=9, depth=depth)
    c = api.Cascador(path, double=True)
    h = oracle.load(path, True)
    assert (c.T, c.K, c.L, c.depth) == (2, 50, 9, depth) == oracle.dims(h)
    a, b = tmp_path / "a", tmp_path / "b"
    c.save_f32(str(a)); oracle.save_f32(h, str(b))
    assert a.read_bytes() == b.read_bytes()
    c.close(); oracle.release(h)


def test_serialiser_flags_stage_field_and_double_flavour(tmp_path):
    """c/jda.c:662-665 writes the header's stage field as T + 1, which cascador.cpp:138 refuses; jdaB200SerializeTo
    can write T and the double flavour the C++ loader reads."""
    import struct
    c = api.Cascador(SHIPPED_F32, double=False)
    plain, fixed, wide = tmp_path / "p", tmp_path / "f", tmp_path / "w"
    c.save(str(plain), 0)
    assert plain.read_bytes() == open(SHIPPED_F32, "rb").read()                  # flags 0 = jdaCascadorSerializeTo
    c.save(str(fixed), api.SAVE_STAGE_T)
    fb, pb = fixed.read_bytes(), plain.read_bytes()
    assert struct.unpack("<7i", pb[:28]) == (0, 5, 540, 27, 4, 6, -1)
    assert struct.unpack("<7i", fb[:28]) == (0, 5, 540, 27, 4, 5, -1) and fb[28:] == pb[28:]
    c.save(str(wide), api.SAVE_STAGE_T | api.SAVE_DOUBLE)
    assert wide.stat().st_size == 10476464                                     # the shipped double model's size
    ref_wide = synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "rw"))
    assert wide.read_bytes() == open(ref_wide, "rb").read()
    c2 = api.Cascador(str(wide), double=True)                                  # and back: exact round trip
    back = tmp_path / "b"
    c2.save_f32(str(back))
    assert back.read_bytes() == pb
    c.close(); c2.close()


@pytest.mark.parametrize("w,h,scale,mn,mx", [(640, 480, 1.25, 24, -1), (640, 480, 1.25, 24, 192),
                                              (640, 480, 1.2, 24, -1), (640, 480, 1.25, 40, -1),
                                              (1920, 1080, 1.25, 24, 768), (450, 333, 1.15, 28, 200),
                                              (30, 27, 1.25, 24, -1), (23, 100, 1.25, 24, -1),
                                              (640, 480, 1.0, 24, -1), (640, 480, 0.9, 24, -1),
                                              (640, 480, 1.5, 100, 50), (24, 24, 1.25, 24, -1)])
def test_levels_and_window_counts_match_oracle(oracle, w, h, scale, mn, mx):
    assert api.levels(w, h, scale, mn, mx) == (oracle.levels(w, h, scale, mn, mx) if min(w, h) >= 24 else [])
    assert api.count_windows(w, h, scale, mn, mx) == oracle.count_windows(w, h, scale, mn, mx)


def test_nms_matches_oracle_on_random_boxes(oracle):
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 17, 300):
        boxes = np.stack([rng.integers(0, 200, n), rng.integers(0, 200, n), rng.integers(24, 120, n)], 1).astype(np.int32)
        scores = rng.normal(0, 1, n).astype(np.float32)
        scores[rng.integers(0, max(n, 1), n // 3)] = 0.5   # ties exercise the strict-< exchange sort
        np.testing.assert_array_equal(api.nms(boxes, scores), oracle.nms(boxes, scores))


@pytest.mark.skipif(api.device_count() > 0, reason="a GPU is visible")
def test_detect_without_gpu_fails_loudly():
    c = api.Cascador(SHIPPED_F32, double=False)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.detect(synth.noise_frame(0, 64, 48))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.detect_batch(synth.make_frames("noise", 2, 64, 48))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.detect_mixed([synth.noise_frame(0, 64, 48), synth.noise_frame(1, 50, 70)])
    assert c.detect_mixed([]) == []          # nothing to do: no device needed, no failure
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.detect_batch(synth.make_frames("noise", 2, 64, 48), flat=True)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.detect_cpp(synth.noise_frame(0, 64, 48))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.trace_cpp(synth.noise_frame(0, 64, 48))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        c.trace(synth.noise_frame(0, 64, 48))
    c.close()


@pytest.mark.parametrize("w,h,mx", [(640, 480, 192), (640, 480, -1), (1920, 1080, 768), (60, 300, -1), (451, 333, -1)])
@pytest.mark.parametrize("latency", [False, True])
def test_tile_plan_invariants(w, h, mx, latency):
    """every planned shared-memory tile fits the pooled scratch and TMA's box rules"""
    plan = api.describe_plan(w, h, 1.25, 24, mx, latency=latency)
    assert [p["win"] for p in plan] == api.levels(w, h, 1.25, 24, mx)
    for p in plan:
        assert p["step"] == int(np.float32(p["win"]) * np.float32(0.1))
        assert p["tw"] * p["th"] <= 512 and p["tw"] in (2, 4, 8, 16, 32)
        if p["smem"]:
            assert p["box_w"] % 16 == 0 and p["box_w"] <= 256 and p["box_h"] <= 256
            assert p["span"] in (1, 2, 4, 12) and p["box_w"] * p["box_h"] <= min(8192 * p["span"], 65536)
            slack = 0 if (p["tw"] * p["step"]) % 16 == 0 else 15
            assert p["box_w"] >= (p["tw"] - 1) * p["step"] + p["win"] + slack
            assert p["box_h"] == (p["th"] - 1) * p["step"] + p["win"]
            assert (p["win"] - 1) * p["box_w"] + p["win"] - 1 < 65536
    assert any(p["smem"] for p in plan)


def test_cli_convert_and_info(tmp_path, capsys):
    from jda_b200.__main__ import main
    wide = synth.widen_f32_model(SHIPPED_F32, str(tmp_path / "wide.model"))
    out = tmp_path / "f32.model"
    assert main(["convert", wide, str(out)]) == 0
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()
    assert main(["info", str(out), "--float", "--size", "640x480", "--max-size", "192"]) == 0
    txt = capsys.readouterr().out
    assert "169236 candidate windows" in txt and "T=5 K=540 landmarks=27" in txt


REF_HEADER_DIR = "/root/reference/c"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_HEADER_DIR, "jda.h")), reason="reference header not present on this box")
def test_c_program_built_against_the_reference_header_runs_on_our_library(tmp_path):
    """drop-in at the link level: a C99 program compiled against the REFERENCE's own c/jda.h (not ours) links to
    libjda_b200.so and its host-only calls behave -- load, serialise (byte-identical), release, NULL safety"""
    import subprocess
    src = tmp_path / "prog.c"
    src.write_text(r'''
#include <stdio.h>
#include "jda.h"
int main(int argc, char **argv) {
  void *c = jdaCascadorCreateFloat(argv[1]);
  if (!c) return 2;
  jdaCascadorSerializeTo(c, argv[2]);
  jdaCascadorRelease(c);
  jdaCascadorRelease(NULL);
  if (jdaCascadorCreateDouble("/nonexistent/model") != NULL) return 3;
  jdaResult r; r.n = 0; r.landmark_n = 27; r.bboxes = NULL; r.shapes = NULL; r.scores = NULL;
  jdaResultRelease(r);
  printf("sizeof(jdaResult)=%d\n", (int)sizeof(jdaResult));
  return 0;
}
''')
    exe = tmp_path / "prog"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", REF_HEADER_DIR, str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(api.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = tmp_path / "rt.model"
    p = subprocess.run([str(exe), SHIPPED_F32, str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "sizeof(jdaResult)=32" in p.stdout
    assert out.read_bytes() == open(SHIPPED_F32, "rb").read()


def test_our_header_is_plain_c(tmp_path):
    """include/jda_b200.h compiles as C99 and as C++ with nothing but the standard headers"""
    import subprocess
    inc = os.path.join(ROOT, "include")
    for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
        subprocess.run(["gcc", "-x", lang, std, "-Wall", "-Werror", "-fsyntax-only", "-I", inc,
                        os.path.join(inc, "jda_b200.h")], check=True)


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The structs of include/jda_b200.h as the C compiler lays them out against their ctypes mirrors in jda_b200/api.py
    (tests and bench call through those): size and the offset of every field."""
    import ctypes
    import subprocess
    pairs = {"jdaB200CppParams": api.CppParams, "jdaB200Stats": api.Stats, "jdaB200FlatResult": api.FlatResult,
             "jdaB200ResultF64": api.ResultF64, "jdaB200Batch": api.Batch, "jdaB200Frame": api.Frame, "jdaResult": api._Result}
    lines = []
    for cname, cls in pairs.items():
        lines.append('printf("%s size %%d\\n", (int)sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%d\\n", (int)offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jda_b200.h"\nint main(void) {\n' + "\n".join(lines) +
                   "\nreturn 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
