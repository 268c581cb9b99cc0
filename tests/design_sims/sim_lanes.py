"""Lane-efficiency model of k2_scan scheduling variants, driven by the oracle's per-window reject
positions.  Offline design tool (CPU only): counts warp-level cart iterations ("packet-carts")."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from jda_b200 import synth

MODEL = "tests/golden/jda_shipped_f32.model"
SCHED = [4, 8, 12, 16, 24, 32, 48, 64, 96, 128, 160, 192, 256, 320, 384, 448, 540]
TILE_BYTES, LIST_CAP = 8192, 512

def plan(win, step, nx, ny, tile_bytes=TILE_BYTES, cap=LIST_CAP):
    best = None
    for tl in (5, 4, 3):
        tw = 1 << tl
        bw = ((tw - 1) * step + win + 15) & ~15
        if bw > 256: continue
        bh_max = min(256, tile_bytes // bw)
        if bh_max < win: continue
        th = min((bh_max - win) // step + 1, cap // tw, max(1, ny))
        windows = min(tw, nx) * th
        if best is None or windows > best[0]: best = (windows, tw, th, True)
    if best is None or best[0] < 64: return (cap, 32, cap // 32, False)
    return best

def levels(w, h, mx):
    o = pyoracle.Oracle(); return o.levels(w, h, 1.25, 24, mx)

def tiles_of_frame(tn, w, h, mx, tile_bytes=TILE_BYTES, cap=LIST_CAP):
    """yield (use_smem, death array per tile in dense order incl. masked lanes = 0)"""
    off = 0
    for win in levels(w, h, mx):
        step = int(np.float32(win) * np.float32(0.1))
        nx, ny = (w - win) // step + 1, (h - win) // step + 1
        d = np.minimum(tn[off:off + nx * ny], 540).reshape(ny, nx); off += nx * ny
        _, tw, th, smem = plan(win, step, nx, ny, tile_bytes, cap)
        for y0 in range(0, ny, th):
            for x0 in range(0, nx, tw):
                blk = np.zeros((min(th, ny - y0), tw), np.int32)
                sub = d[y0:y0 + th, x0:x0 + tw]
                blk[:, :sub.shape[1]] = sub
                yield smem, win, blk.reshape(-1)

def cost_current(tiles, nw, rem1=False, sched=SCHED):
    """returns (packet-cart iterations weighted by packets per iteration, ideal)"""
    tot = 0.0; ideal = 0.0
    for smem, win, d in tiles:
        ideal += d.sum() / 32.0
        cur = d.copy()          # deaths; dense order
        cart = 0
        alive_idx = np.arange(len(cur))
        for cend in sched:
            n = len(alive_idx)
            if n == 0: break
            dd = cur[alive_idx]
            G = 32 * nw
            for b in range(0, n, G):
                g = dd[b:b + G]
                its = min(int(g.max()), cend) - cart
                its = max(its, 0)
                width = nw
                if rem1 and len(g) <= 32: width = 1
                elif rem1: width = (len(g) + 31) // 32
                tot += its * width
            alive_idx = alive_idx[dd > cend] if cend < 540 else alive_idx[:0]
            cart = cend
    return tot, ideal

def cost_pooled(tiles, nw, sched=SCHED):
    """ideal block-level pooling: per phase boundary, survivors from all tiles pool into full groups;
    in-phase deaths still waste lanes."""
    alld = np.concatenate([d for _, _, d in tiles]); alld = alld[alld > 0]
    tot = 0.0; cart = 0
    cur = alld
    for cend in sched:
        n = len(cur)
        if n == 0: break
        # groups formed arbitrarily: expected iterations per group = E[max over 32 of min(d,cend)-cart]
        rng = np.random.default_rng(0); p = rng.permutation(n)
        dd = np.minimum(cur[p], cend) - cart
        pad = (-n) % 32
        dd = np.concatenate([dd, np.zeros(pad, dd.dtype)]).reshape(-1, 32)
        tot += dd.max(1).sum()
        cur = cur[cur > cend]; cart = cend
    return tot, alld.sum() / 32.0

if __name__ == "__main__":
    o = pyoracle.Oracle(); h = o.load(MODEL, False)
    for name, img in [("noise", synth.noise_frame(0)), ("blur6", synth.blur_frame(1)), ("facemix", synth.facemix_frame(3))]:
        tn, ts, _ = o.trace(h, img, max_size=192, t_limit=1)
        T = list(tiles_of_frame(tn, 640, 480, 192))
        print(name, "carts/window %.1f" % (np.minimum(tn, 540).mean()))
        for nw in (1, 2, 4):
            c, i = cost_current(T, nw); c1, _ = cost_current(T, nw, rem1=True)
            print("  current NW=%d: %.2fx ideal ; with NW=1 remainder: %.2fx" % (nw, c / i, c1 / i))
        T2 = list(tiles_of_frame(tn, 640, 480, 192, 16384, 1024))
        c, i = cost_current(T2, 2, rem1=True); print("  16KB/1024-window tiles NW=2 rem1: %.2fx" % (c / i))
        c, i = cost_pooled(T, 1); print("  pooled (block-level buckets): %.2fx" % (c / i))
        fine = [2,4,6,8,10,12,14,16,20,24,28,32,40,48,56,64,80,96,112,128,160,192,224,256,320,384,448,540]
        c, i = cost_pooled(T, 1, fine); print("  pooled, finer schedule: %.2fx" % (c / i))
