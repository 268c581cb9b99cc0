"""Command-line glue around libjda_b200.so (the reference's src/test.cpp / c/main.cpp
are GUI/OpenCV demos and are not rebuilt).

  python -m jda_b200 info    MODEL [--float] [--size WxH] [--max-size N]
  python -m jda_b200 convert MODEL_DOUBLE OUT_FLOAT32          (jdaCascadorSerializeTo, c/jda.c:644-716)
  python -m jda_b200 convert MODEL OUT [--float] --cpp-loadable  (double flavour, header stage field T: what the
                                                                  C++ loader accepts, cascador.cpp:126-164)
  python -m jda_b200 detect  MODEL IMAGE... [--float] [--fddb-out FILE] [--scale S --min-size N --max-size N --th T]
  python -m jda_b200 detect  MODEL IMAGE... --cpp [--fddb-min 20 --fddb-step 5 --fddb-scale 1.2 --overlap 0.3 --no-nms]

  python -m jda_b200 fddb    MODEL FDDB_DIR [--float] [--c-api] [--folds 1-10]

`fddb` is the reference's FDDB runner (src/test.cpp:73-235) without the drawing: for every fold it reads
FDDB_DIR/FDDB-folds/FDDB-fold-XX.txt, opens FDDB_DIR/images/<path>.jpg (unreadable images are skipped like there),
converts BGR -> gray, runs the detector over all images of the fold in one batch and writes
FDDB_DIR/result/fold-XX-out.txt in the evaluator's format.  Default detector: JoinCascador::Detect (double precision,
config.json's fddb settings) as in the reference; --c-api uses jdaDetect's semantics instead.

`--cpp` runs the reference's double-precision C++ detector (JoinCascador::Detect, fddb.method = 1 -- what src/test.cpp's
FDDB runner itself calls) with the fddb.* settings of config.json instead of the C library's jdaDetect.

`detect` reads gray images (.npy u8 arrays, binary .pgm, or anything cv2 can open when cv2 is present) and, with
--fddb-out, writes the result-file format of the reference's FDDB runner (src/test.cpp:153,163):
    <path>\\n<n>\\n<x> <y> <w> <h> <score>\\n ...
"""
import argparse
import os
import sys

import numpy as np

from . import api


def read_gray(path):
    if path.endswith(".npy"):
        a = np.load(path)
    elif path.endswith(".pgm"):
        with open(path, "rb") as f:
            tok = []
            while len(tok) < 4:
                line = f.readline()
                if not line.startswith(b"#"):
                    tok += line.split()
            assert tok[0] == b"P5" and int(tok[3]) == 255, "only 8-bit binary PGM"
            w, h = int(tok[1]), int(tok[2])
            a = np.frombuffer(f.read(w * h), np.uint8).reshape(h, w)
    else:
        import cv2
        a = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
        if a is None:
            raise SystemExit("cannot read " + path)
    if a.ndim == 3:
        a = a.mean(axis=2)
    return np.ascontiguousarray(a, np.uint8)


def run_fddb(c, a):
    """src/test.cpp:96-224 without the drawing"""
    import cv2
    lo, _, hi = a.folds.partition("-")
    os.makedirs(os.path.join(a.fddb_dir, "result"), exist_ok=True)
    total = 0
    for i in range(int(lo), int(hi or lo) + 1):
        fold = os.path.join(a.fddb_dir, "FDDB-folds", "FDDB-fold-%02d.txt" % i)
        names, frames = [], []
        for path in open(fold).read().split():
            img = cv2.imread(os.path.join(a.fddb_dir, "images", path + ".jpg"))          # test.cpp:126
            if img is None:
                print("Can not open %s, Skip it" % path)
                continue
            names.append(path)
            frames.append(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))                         # test.cpp:132
        if a.c_api:
            res = [(np.column_stack([b, b[:, 2]]) if len(b) else np.zeros((0, 4), np.int32), s, p)
                   for b, s, p in c.detect_many(frames)]
        else:
            res = c.detect_cpp_many(frames, minimum_size=a.fddb_min, step=a.fddb_step, scale=a.fddb_scale,
                                    overlap=a.overlap, nms=not a.no_nms, similarity=a.similarity_transform,
                                    shift=tuple(a.shift))
        with open(os.path.join(a.fddb_dir, "result", "fold-%02d-out.txt" % i), "w") as out:
            for path, (rects, scores, _) in zip(names, res):
                out.write("%s\n%d\n" % (path, len(scores)))                            # test.cpp:153
                for r, sc in zip(rects, scores):
                    out.write("%d %d %d %d %f\n" % (r[0], r[1], r[2], r[3], sc))        # test.cpp:163
        total += len(names)
        print("fold %02d: %d images, %d detections" % (i, len(names), sum(len(r[1]) for r in res)))
    return 0


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m jda_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("info")
    p.add_argument("model"); p.add_argument("--float", action="store_true")
    p.add_argument("--size", default="640x480"); p.add_argument("--max-size", type=int, default=-1)
    p = sub.add_parser("convert")
    p.add_argument("model"); p.add_argument("out")
    p.add_argument("--float", action="store_true", help="the input is a float32-flavour file")
    p.add_argument("--cpp-loadable", action="store_true",
                   help="write the double flavour with the header's stage field = T (cascador.cpp:138), not the C "
                        "serialiser's T + 1 (c/jda.c:662-665)")
    p = sub.add_parser("detect")
    p.add_argument("model"); p.add_argument("images", nargs="+"); p.add_argument("--float", action="store_true")
    p.add_argument("--fddb-out"); p.add_argument("--scale", type=float, default=1.25)
    p.add_argument("--min-size", type=int, default=24); p.add_argument("--max-size", type=int, default=-1)
    p.add_argument("--th", type=float, default=0.0)
    p.add_argument("--cpp", action="store_true", help="JoinCascador::Detect (double precision, fddb.method 1)")
    p.add_argument("--fddb-min", type=int, default=20); p.add_argument("--fddb-step", type=int, default=5)
    p.add_argument("--fddb-scale", type=float, default=1.2); p.add_argument("--overlap", type=float, default=0.3)
    p.add_argument("--no-nms", action="store_true")
    p.add_argument("--similarity-transform", action="store_true", help="config.json face.similarity_transform (C++ detector)")
    p.add_argument("--shift", type=float, nargs=2, default=(0.0, 0.0), metavar=("X", "Y"),
                   help="initial shift added to the mean shape (face.random_shift draws one per window in the reference; "
                        "src/test.cpp forces 0)")
    p = sub.add_parser("fddb")
    p.add_argument("model"); p.add_argument("fddb_dir"); p.add_argument("--float", action="store_true")
    p.add_argument("--c-api", action="store_true", help="jdaDetect semantics instead of JoinCascador::Detect")
    p.add_argument("--folds", default="1-10")
    p.add_argument("--fddb-min", type=int, default=20); p.add_argument("--fddb-step", type=int, default=5)
    p.add_argument("--fddb-scale", type=float, default=1.2); p.add_argument("--overlap", type=float, default=0.3)
    p.add_argument("--no-nms", action="store_true")
    p.add_argument("--similarity-transform", action="store_true", help="config.json face.similarity_transform (C++ detector)")
    p.add_argument("--shift", type=float, nargs=2, default=(0.0, 0.0), metavar=("X", "Y"),
                   help="initial shift added to the mean shape (face.random_shift draws one per window in the reference; "
                        "src/test.cpp forces 0)")
    a = ap.parse_args(argv)

    if a.cmd == "convert":
        c = api.Cascador(a.model, double=not a.float)
        if a.cpp_loadable:
            c.save(a.out, api.SAVE_STAGE_T | api.SAVE_DOUBLE)
        else:
            c.save_f32(a.out)
        print("wrote %s (%d bytes, %s flavour, T=%d K=%d L=%d depth=%d)" %
              (a.out, os.path.getsize(a.out), "double" if a.cpp_loadable else "float32", c.T, c.K, c.L, c.depth))
        return 0
    c = api.Cascador(a.model, double=not a.float)
    if a.cmd == "info":
        w, h = map(int, a.size.split("x"))
        print("T=%d K=%d landmarks=%d depth=%d" % (c.T, c.K, c.L, c.depth))
        print("%dx%d, max_size %d: %d candidate windows" % (w, h, a.max_size, api.count_windows(w, h, 1.25, 24, a.max_size)))
        for lat in (False, True):
            print("scan plan (%s):" % ("calls of <= 4 frames or <= 2.1e6 candidate windows" if lat else
                                       "batches; up to 8e6 candidate windows the global-memory levels use 32 x 4 tiles"))
            for q in api.describe_plan(w, h, 1.25, 24, a.max_size, latency=lat):
                print("   win %3d step %2d  %4d x %-4d windows  tile %2d x %-2d  box %3d x %-3d  %s" %
                      (q["win"], q["step"], q["nx"], q["ny"], q["tw"], q["th"], q["box_w"], q["box_h"],
                       ("shared memory, %d warp(s) per tile" % q["span"]) if q["smem"] else "global memory"))
        return 0
    if a.cmd == "fddb":
        return run_fddb(c, a)
    frames = [read_gray(p) for p in a.images]
    if a.cpp:
        res = c.detect_cpp_many(frames, minimum_size=a.fddb_min, step=a.fddb_step, scale=a.fddb_scale, overlap=a.overlap,
                                nms=not a.no_nms, similarity=a.similarity_transform, shift=tuple(a.shift))
    else:
        res = c.detect_many(frames, scale=a.scale, min_size=a.min_size, max_size=a.max_size, th=a.th)
    out = open(a.fddb_out, "w") if a.fddb_out else sys.stdout
    for path, (boxes, scores, shapes) in zip(a.images, res):
        name = os.path.splitext(path)[0]
        out.write("%s\n%d\n" % (name, len(scores)))
        for b, s in zip(boxes, scores):
            out.write("%d %d %d %d %f\n" % (b[0], b[1], b[2], b[3] if a.cpp else b[2], s))  # test.cpp:163: "%d %d %d %d %lf"
    if a.fddb_out:
        out.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
