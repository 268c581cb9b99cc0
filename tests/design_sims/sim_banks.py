"""Shared-memory bank-conflict model for the pixel reads of k2_scan: for each level and candidate
(tw, tile pitch) replay the phase/compaction schedule on oracle reject positions and count LDS
wavefronts per pixel read (= max number of distinct 4-byte words per bank over the 32 lanes)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from jda_b200 import synth
from tests.design_sims.sim_lanes import SCHED

def wavefronts(addr):
    words = addr >> 2
    banks = words & 31
    mx = 0
    for b in np.unique(banks):
        mx = max(mx, len(np.unique(words[banks == b])))
    return mx

def level_cost(d2, win, step, tw, th, pitch, rng, nw=4, samples_per_phase=3):
    """d2: [ny, nx] deaths.  returns (sum wavefronts*iterations, sum iterations) over sampled packets."""
    ny, nx = d2.shape
    tot_w = 0.0; tot_i = 0.0
    offs = rng.integers(0, win, (samples_per_phase, 2))   # random pixel offsets inside the window
    for y0 in range(0, ny, th):
        for x0 in range(0, nx, tw):
            sub = d2[y0:y0 + th, x0:x0 + tw]
            hh, ww = sub.shape
            wy, wx = np.mgrid[0:hh, 0:ww]
            base = (wy * step * pitch + wx * step).reshape(-1)
            dd = sub.reshape(-1)
            cart = 0
            idx = np.arange(len(dd))
            for cend in SCHED:
                if len(idx) == 0: break
                if len(idx) <= 15: break     # straggler mode from here
                for b in range(0, len(idx), 32):
                    g = idx[b:b + 32]
                    its = min(int(dd[g].max()), cend) - cart
                    if its <= 0: continue
                    # lanes alive at the START of the phase all issue loads (dead ones read the origin)
                    w = np.mean([wavefronts(base[g] + oy * pitch + ox) for oy, ox in offs])
                    tot_w += w * its; tot_i += its
                idx = idx[dd[idx] > cend]
                cart = cend
    return tot_w, tot_i

if __name__ == "__main__":
    o = pyoracle.Oracle(); h = o.load("tests/golden/jda_shipped_f32.model", False)
    rng = np.random.default_rng(0)
    frames = [synth.noise_frame(0), synth.blur_frame(1)]
    traces = [o.trace(h, f, max_size=192, t_limit=1)[0] for f in frames]
    wins = o.levels(640, 480, 1.25, 24, 192)
    off = 0
    for win in wins:
        step = int(np.float32(win) * np.float32(0.1))
        nx, ny = (640 - win) // step + 1, (480 - win) // step + 1
        if win > 57: break
        print("win", win, "step", step)
        for tw in (32, 16, 8):
            need = (tw - 1) * step + win
            base_p = (need + 15) & ~15
            for pitch in range(base_p, min(base_p + 96, 257), 16):
                rows = min(256, 8192 // pitch)
                if rows < win: continue
                th = min((rows - win) // step + 1, 512 // tw)
                if tw * th < 64: continue
                tw_, ti_ = 0, 0
                for tn in traces:
                    d2 = np.minimum(tn[off:off + nx * ny], 540).reshape(ny, nx)
                    a, b = level_cost(d2, win, step, tw, th, pitch, rng)
                    tw_ += a; ti_ += b
                R = (step * pitch // 4) % 32
                print("   tw %2d pitch %3d th %2d (windows %3d, R=%2d): %.2f wavefronts / pixel LDS" % (tw, pitch, th, tw * th, R, tw_ / ti_))
        off += nx * ny
