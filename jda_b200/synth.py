"""Synthetic inputs for the detect path: frames (SURVEY.md 8(d) distributions) and
model files in the reference's binary layout (README.md:84-111, c/jda.c:486-561).

Pure numpy/scipy; deterministic per seed.  Used by tests/ and bench.py.
"""
import os
import struct

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def noise_frame(seed, w=640, h=480):
    """uniform u8 noise -- ~99 carts/window with the shipped model."""
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)


def blur_frame(seed, w=640, h=480, sigma=6.0):
    """Gaussian-blurred noise, min-max rescaled to 0..255 -- natural-image-like (~37 carts/window)."""
    from scipy.ndimage import gaussian_filter
    f = np.random.default_rng(seed).integers(0, 256, (h, w)).astype(np.float32)
    if sigma > 0:
        f = gaussian_filter(f, sigma, mode="reflect")
    lo, hi = float(f.min()), float(f.max())
    return ((f - lo) * (255.0 / max(hi - lo, 1e-6))).astype(np.uint8)


_face_cache = {}


def face_crop():
    """222x216 gray face crop (fixture derived from the reference's model/jda-27.png
    by tests/golden/make_fixtures.py)."""
    if "f" not in _face_cache:
        _face_cache["f"] = np.load(os.path.join(GOLDEN, "face_222x216.npy"))
    return _face_cache["f"]


def _resize_area(img, dw, dh):
    """box-filter resize in float64 (separable area weights); deterministic numpy only."""
    def weights(s, d):
        m = np.zeros((d, s))
        r = s / d
        for i in range(d):
            a, b = i * r, (i + 1) * r
            j0, j1 = int(np.floor(a)), min(int(np.ceil(b)), s)
            for j in range(j0, j1):
                m[i, j] = max(0.0, min(b, j + 1) - max(a, j))
            m[i] /= m[i].sum()
        return m
    h, w = img.shape
    out = weights(h, dh) @ img.astype(np.float64) @ weights(w, dw).T
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def face_canvas():
    """SURVEY.md section 4 known-answer canvas: gray 100 VGA, face at 222x216 @ (60,40)
    and 111x108 @ (400,300).  Both crops are stored fixtures (cv2 INTER_AREA here)."""
    c = np.full((480, 640), 100, np.uint8)
    big = face_crop()
    small = np.load(os.path.join(GOLDEN, "face_111x108.npy"))
    c[40:40 + 216, 60:60 + 222] = big
    c[300:300 + 108, 400:400 + 111] = small
    return c


def facemix_frame(seed, w=640, h=480):
    """blur6 background + 0..4 pasted faces at seeded positions / sizes 40..300 px."""
    f = blur_frame(seed, w, h).copy()
    rng = np.random.default_rng(seed + 7919)
    face = face_crop()
    for _ in range(int(rng.integers(0, 5))):
        s = int(rng.integers(40, min(300, w, h) + 1))
        fw, fh = s, max(8, int(round(s * face.shape[0] / face.shape[1])))
        if fw > w or fh > h:
            continue
        x = int(rng.integers(0, w - fw + 1))
        y = int(rng.integers(0, h - fh + 1))
        f[y:y + fh, x:x + fw] = _resize_area(face, fw, fh)
    return f


MINING_SIGMAS = (0.0, 1.0, 2.0, 4.0, 6.0)


def mining_background(seed, w=640, h=480):
    """a face-free background, noise blurred with sigma in {0, 1, 2, 4, 6} by seed mod 5."""
    return blur_frame(seed, w, h, sigma=MINING_SIGMAS[seed % 5])


DISTRIBUTIONS = {"noise": noise_frame, "blur6": blur_frame, "facemix": facemix_frame}


def make_frames(dist, n, w=640, h=480, seed0=0):
    """[n, h, w] u8 frames; dist in noise|blur6|facemix|mix (mix = round-robin of the three)."""
    out = np.empty((n, h, w), np.uint8)
    names = ["noise", "blur6", "facemix"]
    for i in range(n):
        d = names[i % 3] if dist == "mix" else dist
        out[i] = DISTRIBUTIONS[d](seed0 + i, w, h)
    return out


def fddb_shape(seed):
    """FDDB-like frame size: longest side 450, short side in [229,450], coin-flip orientation."""
    rng = np.random.default_rng(seed + 104729)
    short = int(rng.integers(229, 451))
    return (450, short) if rng.integers(0, 2) else (short, 450)  # (w, h)


# ------------------------------------------------------------------ models

def write_model(path, seed=0, T=5, K=540, L=27, depth=4, double=True, scales=(0,),
                mode="reject", norm_every=270, w_amp=2e-4, coord_max=None, scales_by_stage=None,
                header_stage=None):
    """Random model in the reference's file layout.

    mode: 'passall' -> every cart threshold -1e30 (all windows run T*K carts + T regressions)
          'reject'  -> thresholds on a slowly falling ramp so windows die at varied depths
    scales: allowed node.scale values; with 1/2 present pass coord_max (~0.45) so sampled
            coordinates stay inside the region where the reference's h/q indexing is defined.
    norm_every: carts k with (k+1) % norm_every == 0 get a non-trivial (mean, std).
    scales_by_stage: {stage: allowed scales} overriding `scales` for those stages, e.g. {0: (0,)} keeps
            stage 0 on the o plane (the LUT scan serves it) while later stages sample the h / q planes.
    header_stage: value of the header's current_stage_idx field (default T, what the C++ trainer writes
            for a finished model; the C serialiser writes T + 1, c/jda.c:662-665).
    """
    rng = np.random.default_rng(seed)
    nl = 1 << (depth - 1)
    nn = nl - 1
    real = "<f8" if double else "<f4"
    D = 2 * L
    mean = rng.uniform(0.25, 0.75, D)
    if coord_max is not None:
        mean = rng.uniform(0.12, coord_max - 0.1, D)
    chunks = [struct.pack("<7i", 0, T, K, L, depth, T if header_stage is None else header_stage, -1),
              mean.astype(real).tobytes()]
    for t in range(T):
        st_scales = (scales_by_stage or {}).get(t, scales)
        for k in range(K):
            for i in range(nn):
                sc = int(rng.choice(st_scales))
                l1, l2 = int(rng.integers(0, L)), int(rng.integers(0, L))
                off = rng.uniform(-0.12, 0.12, 4)
                if coord_max is not None:
                    off = rng.uniform(-0.08, 0.08, 4)
                th = int(rng.integers(-60, 61))
                chunks.append(struct.pack("<3i", sc, l1, l2) + off.astype(real).tobytes()
                              + struct.pack("<i", th))
            chunks.append(rng.normal(0, 1, nl).astype(real).tobytes())
            if mode == "passall":
                cth = -1e30
            else:
                cth = -1.0 - 0.05 * np.sqrt(t * K + k + 1.0) + rng.normal(0, 0.3)
            if norm_every and (k + 1) % norm_every == 0:
                mu, sd = rng.normal(0, 0.5), rng.uniform(0.7, 1.6)
            else:
                mu, sd = 0.0, 1.0
            chunks.append(np.array([cth, mu, sd]).astype(real).tobytes())
        chunks.append(rng.normal(0, w_amp, (K * nl, D)).astype(real).tobytes())
    chunks.append(struct.pack("<i", 0))
    with open(path, "wb") as f:
        for c in chunks:
            f.write(c)
    return path


def widen_f32_model(src, dst, T=5, K=540, L=27, depth=4):
    """float32-flavour file -> double-flavour file with identical values (f32 -> f64 is exact),
    header stage field normalised to T (the float writer stores T+1, c/jda.c:662-665)."""
    nl = 1 << (depth - 1)
    nn = nl - 1
    D = 2 * L
    b = open(src, "rb").read()
    o = 0
    out = []

    def take(n):
        nonlocal o
        s = b[o:o + n]
        o += n
        return s

    hdr = list(struct.unpack("<7i", take(28)))
    hdr[5] = T
    out.append(struct.pack("<7i", *hdr))

    def reals(n):
        return np.frombuffer(take(4 * n), "<f4").astype("<f8").tobytes()

    out.append(reals(D))
    for t in range(T):
        for k in range(K):
            for i in range(nn):
                out.append(take(12))
                out.append(reals(4))
                out.append(take(4))
            out.append(reals(nl + 3))
        out.append(reals(K * nl * D))
    out.append(take(4))
    assert o == len(b), (o, len(b))
    with open(dst, "wb") as f:
        for c in out:
            f.write(c)
    return dst
