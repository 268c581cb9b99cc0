"""Exact shared-memory wavefront model of k2_scan's pixel reads.

For one frame and one pyramid level it replays the scan kernel's phases on the oracle's reject positions
and on the real tree paths (recomputed here with numpy from the model file and the frame), and counts,
for every LDS.U8 of every packet-cart, the wavefronts the access takes: max over the 32 banks of the
number of distinct 4-byte words the lanes touch in that bank.  Output: average wavefronts per pixel load
by tree depth (root / level 1 / level 2) and by phase, for a given tile shape and pitch -- the number
ncu reports as 2.36 per LDS.U8 (profiles/r1b_hotspots_k2_k3.txt, lines 233/234).

Design tool only (CPU, oracle-driven): python -m tests.design_sims.sim_exact_banks [win] [frame kind]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from jda_b200 import synth, api

MODEL = "tests/golden/jda_shipped_f32.model"
SCHED = [8, 16, 24, 32, 40, 48, 64, 80, 96, 128, 160, 192, 256, 320, 384, 448, 540]   # api.cu: ctx_init (round-2 default)
STRAG_MAX = 32                                                                        # kernels.cuh: K2_STRAG_MAX
K = 540


def load_stage0(path=MODEL):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, np.int32, 7)
    assert tuple(hdr[1:5]) == (5, 540, 27, 4)
    mean = np.frombuffer(raw, np.float32, 54, 28)
    cart = np.frombuffer(raw, np.uint8, K * 268, 28 + 216).reshape(K, 268)
    nodes = cart[:, :224].reshape(K, 7, 32)
    lm1 = nodes[:, :, 4:8].copy().view(np.int32)[..., 0]
    lm2 = nodes[:, :, 8:12].copy().view(np.int32)[..., 0]
    off = nodes[:, :, 12:28].copy().view(np.float32)          # o1x o1y o2x o2y
    th = nodes[:, :, 28:32].copy().view(np.int32)[..., 0]
    return mean, lm1, lm2, off, th


def node_xy(mean, lm1, lm2, off, win):
    """integer pixel coordinates of both points of every stage-0 node (c/jda.c:373-389 at the mean shape)"""
    f = np.float32
    v = np.stack([mean[2 * lm1] + off[..., 0], mean[2 * lm1 + 1] + off[..., 1],
                  mean[2 * lm2] + off[..., 2], mean[2 * lm2 + 1] + off[..., 3]], -1).astype(f)
    c = (v * f(win)).astype(np.int32)           # truncation toward zero like (int)
    return np.clip(c, 0, win - 1)               # [K, 7, 4] = x1 y1 x2 y2


def tree_paths(img, win, step, xy, th):
    """node index visited at depth 1 and 2 for every window of the level and every cart: [K, ny*nx] each"""
    H, W = img.shape
    nx, ny = (W - win) // step + 1, (H - win) // step + 1
    ys, xs = np.mgrid[0:ny, 0:nx]
    base = (ys * step * W + xs * step).reshape(-1)
    flat = img.reshape(-1).astype(np.int32)
    n1 = np.empty((K, nx * ny), np.uint8)
    n2 = np.empty((K, nx * ny), np.uint8)
    o1 = xy[..., 1] * W + xy[..., 0]
    o2 = xy[..., 3] * W + xy[..., 2]
    for k in range(K):
        def test(node):   # node: scalar or per-window array
            a = flat[base + o1[k][node]] - flat[base + o2[k][node]]
            return np.where(a <= th[k][node], 1, 2)
        i1 = test(0)
        i2 = 2 * i1 + test(i1)
        n1[k], n2[k] = i1, i2
    return n1, n2, nx, ny


def wavefronts(addr):
    """addr [P, 32] byte addresses -> [P] wavefronts (distinct words per bank, max over banks)"""
    words = np.sort(addr >> 2, axis=1)
    uniq = np.ones_like(words, bool)
    uniq[:, 1:] = words[:, 1:] != words[:, :-1]
    P = addr.shape[0]
    key = (np.arange(P)[:, None] * 32 + (words & 31)).reshape(-1)
    cnt = np.bincount(key, weights=uniq.reshape(-1), minlength=P * 32).reshape(P, 32)
    return cnt.max(axis=1)


def simulate(img, deaths, win, step, tw, th_rows, pitch, xy, n1, n2, nx, ny, nw=4, order="row", max_tiles=None, rng=None,
             planes=False, bank_order=False):
    """returns dict depth -> [sum wavefronts, loads] and per-phase totals.
    planes=True: the de-interleaved tile layout (round 2): pixel (x, y) of the tile lives in plane (y % step, x % step) at
    (y // step, x // step), so a window's base is wy * bx + wx and neighbouring windows are ONE byte apart; `pitch` is bx."""
    if planes:
        halo = (win - 1) // step
        by = th_rows + halo
        psz = pitch * by
        o1 = ((xy[..., 1] % step) * step + xy[..., 0] % step) * psz + (xy[..., 1] // step) * pitch + xy[..., 0] // step
        o2 = ((xy[..., 3] % step) * step + xy[..., 2] % step) * psz + (xy[..., 3] // step) * pitch + xy[..., 2] // step
    else:
        o1 = xy[..., 1] * pitch + xy[..., 0]     # tile-format offsets with this pitch
        o2 = xy[..., 3] * pitch + xy[..., 2]
    tot = np.zeros((3, 2))
    per_phase = np.zeros((len(SCHED), 2))
    d2 = np.minimum(deaths, K).reshape(ny, nx)
    tiles = [(y0, x0) for y0 in range(0, ny, th_rows) for x0 in range(0, nx, tw)]
    if max_tiles and len(tiles) > max_tiles:
        tiles = [tiles[i] for i in (rng or np.random.default_rng(0)).choice(len(tiles), max_tiles, replace=False)]
    for y0, x0 in tiles:
        hh, ww = min(th_rows, ny - y0), min(tw, nx - x0)
        wy, wx = np.mgrid[0:hh, 0:tw]
        valid = (wx < ww).reshape(-1)
        gidx = ((y0 + wy) * nx + np.minimum(x0 + wx, nx - 1)).reshape(-1)   # index into the level arrays
        tbase = (wy * pitch + wx).reshape(-1) if planes else (wy * step * pitch + wx * step).reshape(-1)
        dd = np.where(valid, d2.reshape(-1)[gidx], 0)
        # lane 'window 0' stand-in for masked / dead lanes: the tile origin
        g0, b0 = gidx[0], 0
        alive = np.nonzero(valid)[0] if False else np.arange(hh * tw)        # phase 0 enumerates densely
        if order.startswith("block"):   # block-major enumeration: blocks of bw x bh windows, row-major inside a block
            bw_, bh_ = (int(v) for v in order[5:].split("x"))
            key = ((wy // bh_) * ((tw + bw_ - 1) // bw_) + wx // bw_) * (bw_ * bh_) + (wy % bh_) * bw_ + wx % bw_
            alive = np.argsort(key.reshape(-1)[:hh * tw], kind="stable")
        cart = 0
        for ph, cend in enumerate(SCHED):
            n = len(alive)
            if n == 0:
                break
            if ph > 0 and n <= STRAG_MAX:
                break                                   # straggler mode from here (not modelled)
            if ph > 0 and bank_order:
                # what-if: the compacted list ordered so that every packet of 32 holds windows of distinct bank classes
                # as far as the classes' sizes allow (round-robin over the 32 classes of the windows' base words)
                cls = (tbase[alive] >> 2) & 31
                srt = np.argsort(cls, kind="stable")
                rank = np.empty(n, np.int64)
                starts = np.r_[0, np.cumsum(np.bincount(cls, minlength=32))[:-1]]
                rank[srt] = np.arange(n) - starts[cls[srt]]
                alive = alive[np.lexsort((cls, rank))]
            lst = alive
            npk = (n + 31) // 32
            W_ = np.full(npk * 32, -1, np.int64)
            W_[:n] = lst
            act = W_ >= 0
            if ph == 0:
                act &= np.where(W_ >= 0, valid[np.maximum(W_, 0)], False)
            G = np.where(act, gidx[np.maximum(W_, 0)], g0).reshape(npk, 32)
            B = np.where(act, tbase[np.maximum(W_, 0)], b0).reshape(npk, 32)
            D = np.where(act, dd[np.maximum(W_, 0)], 0).reshape(npk, 32)
            # groups of nw packets share the loop trip count (the __any_sync is per group)
            grp = np.arange(npk) // nw
            # remainder: the kernel narrows the last group; model: trip count by the packets actually present
            gmax = np.zeros(grp.max() + 1, np.int64)
            np.maximum.at(gmax, grp, D.max(axis=1))
            its_end = np.minimum(gmax[grp], cend)     # per packet: last cart index (exclusive)
            for k in range(cart, cend):
                run = its_end > k
                if not run.any():
                    break
                Gk, Bk = G[run], B[run]
                i1 = n1[k][Gk]
                i2 = n2[k][Gk]
                for depth, node in ((0, None), (1, i1), (2, i2)):
                    if node is None:
                        a1 = Bk + o1[k][0]
                        a2 = Bk + o2[k][0]
                    else:
                        a1 = Bk + o1[k][node]
                        a2 = Bk + o2[k][node]
                    w = wavefronts(a1).sum() + wavefronts(a2).sum()
                    tot[depth] += (w, 2 * len(Gk))
                    per_phase[ph] += (w, 2 * len(Gk))
            keep = dd[alive] > cend
            alive = alive[keep]
            cart = cend
    return tot, per_phase


def main():
    win = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    kind = sys.argv[2] if len(sys.argv) > 2 else "facemix"
    img = {"noise": synth.noise_frame(0), "blur6": synth.blur_frame(1), "facemix": synth.facemix_frame(3)}[kind]
    o = pyoracle.Oracle()
    h = o.load(MODEL, False)
    tn, _, _ = o.trace(h, img, max_size=192, t_limit=1)
    plan = api.describe_plan(640, 480, 1.25, 24, 192)
    off = 0
    for p in plan:
        if p["win"] == win:
            break
        off += p["nx"] * p["ny"]
    step = p["step"]
    mean, lm1, lm2, offs, th = load_stage0()
    xy = node_xy(mean, lm1, lm2, offs, win)
    n1, n2, nx, ny = tree_paths(img, win, step, xy, th)
    deaths = tn[off:off + nx * ny]
    print("level win %d step %d: %d windows, avg carts %.1f; shipped plan tw %d th %d pitch %d" %
          (win, step, nx * ny, np.minimum(deaths, K).mean(), p["tw"], p["th"], p["box_w"]))
    cands = [(p["tw"], p["th"], p["box_w"])]
    for a in sys.argv[3:]:
        cands.append(tuple(int(v) for v in a.split(",")))
    halo = (win - 1) // step
    for tw, thr, pitch in cands + [(c[0], c[1], -((c[0] + halo + 15) // 16 * 16)) for c in cands[:1]] + \
            [(32, 16, -((32 + halo + 15) // 16 * 16)), (16, 32, -((16 + halo + 15) // 16 * 16)), (64, 8, -((64 + halo + 15) // 16 * 16))]:
        planes = pitch < 0
        pitch = abs(pitch)
        if planes and step * step * pitch * (thr + halo) > 8192:
            print("(planes tw %d th %d bx %d: %d bytes, does not fit 8 KB)" % (tw, thr, pitch, step * step * pitch * (thr + halo)))
            continue
        bo = os.environ.get("SIM_BANK_ORDER") == "1"
        tot, per = simulate(img, deaths, win, step, tw, thr, pitch, xy, n1, n2, nx, ny, max_tiles=40, planes=planes,
                            order=os.environ.get("SIM_ORDER", "row"), bank_order=bo)
        s = "%s%s tw %2d th %2d pitch %3d: " % ("PLANES" if planes else "raw   ", " bank-ordered packets" if bo else "", tw, thr, pitch)
        s += "  ".join("%s %.2f" % (nm, tot[i, 0] / tot[i, 1]) for i, nm in enumerate(("root", "L1", "L2")))
        s += "  all %.3f wavefronts / LDS.U8" % (tot[:, 0].sum() / tot[:, 1].sum())
        print(s)
        print("     by phase: " + " ".join("%d:%.2f(%.0f%%)" % (SCHED[i], per[i, 0] / per[i, 1], 100 * per[i, 1] / per[:, 1].sum())
                                          for i in range(len(SCHED)) if per[i, 1] > 0))


if __name__ == "__main__":
    main()
