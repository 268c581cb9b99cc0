"""Instruction-cost model of the current k2_scan (phases + 4/2/1-wide remainder groups + straggler mode),
per pyramid level, from oracle reject positions.  Unit: one 32-window packet-cart = 41 instructions."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from jda_b200 import synth, api
from tests.design_sims.sim_lanes import SCHED

def tile_cost(d, nw=4, strag=15):
    """d: deaths (0 = masked lane) in dense order. returns (cost, ideal, parts dict)"""
    ideal = d[d > 0].sum() / 32.0
    cost = 0.0; parts = {"phase": 0.0, "strag": 0.0, "trans": 0.0}
    idx = np.arange(len(d)); cart = 0; first = True
    for cend in SCHED:
        n = len(idx)
        if n == 0: break
        dd = d[idx]
        b = 0
        while b < n:
            rem = n - b
            width = nw if rem >= 32 * nw else (4 if rem > 64 and nw >= 4 else 2 if rem > 32 else 1)
            g = dd[b:b + 32 * width]
            its = max(min(int(g.max()), cend) - cart, 0)
            cost += its * width; parts["phase"] += its * width
            cost += 0.5 * width; parts["trans"] += 0.5 * width       # list read/write + setup ~20 instr per packet
            b += 32 * width
        idx = idx[dd > cend] if cend < 540 else idx[:0]
        cart = cend
        n = len(idx)
        if 0 < n <= strag and cart < 540:
            dd = d[idx]
            alive = dd.copy()
            for k0 in range(cart, 540, 32):
                na = int((alive > k0).sum())
                if na == 0: break
                c = (35 * na + 110) / 41.0
                cost += c; parts["strag"] += c
            break
    return cost, ideal, parts

if __name__ == "__main__":
    o = pyoracle.Oracle(); h = o.load("tests/golden/jda_shipped_f32.model", False)
    plan = api.describe_plan(640, 480, 1.25, 24, 192)
    for name, img in [("noise", synth.noise_frame(0)), ("blur6", synth.blur_frame(1)), ("facemix", synth.facemix_frame(3))]:
        tn, _, _ = o.trace(h, img, max_size=192, t_limit=1)
        off = 0; tot_c = tot_i = 0
        print(name)
        for p in plan:
            nx, ny, tw, th = p["nx"], p["ny"], p["tw"], p["th"]
            d2 = np.minimum(tn[off:off + nx * ny], 540).reshape(ny, nx); off += nx * ny
            c = i = 0; parts = {"phase": 0.0, "strag": 0.0, "trans": 0.0}
            span = p["span"]
            for y0 in range(0, ny, th):
                for x0 in range(0, nx, tw):
                    sub = d2[y0:y0 + th, x0:x0 + tw]
                    rows = sub.shape[0]
                    for g in range(span):       # pooled tiles: rows split between the group's warps
                        r0, r1 = rows * g // span, rows * (g + 1) // span
                        if r1 <= r0: continue
                        blk = np.zeros((r1 - r0, tw), np.int32); blk[:, :sub.shape[1]] = sub[r0:r1]
                        cc, ii, pp = tile_cost(blk.reshape(-1))
                        c += cc; i += ii
                        for k in parts: parts[k] += pp[k]
            tot_c += c; tot_i += i
            print("   win %3d (%4.1f%% of ideal work): cost/ideal %.2f  [phase %.2f strag %.2f trans %.2f]" %
                  (p["win"], 0, c / i, parts["phase"] / i, parts["strag"] / i, parts["trans"] / i))
        print("   total cost/ideal %.2f" % (tot_c / tot_i))
