"""Discrete-event model of the pooled stage-0 scan (k2_pool.cuh policy) for one block.
Time unit = one cart iteration of one warp (latency-bound regime).  Offline design tool."""
import heapq, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pyoracle
from jda_b200 import synth
from tests.design_sims.sim_lanes import tiles_of_frame, SCHED

def simulate(tiles, W=24, R=12, NW=2, sched=SCHED, carry_frac=0.625, cap=192, load_time=3.0, verbose=False):
    G = 32 * NW; carry_min = int(G * carry_frac)
    nph = len(sched)
    tiles = list(tiles); next_tile = 0
    slots = [None] * R        # dict(d=deaths array, cursor, npk, ref)
    buckets = [[] for _ in range(nph + 1)]
    t_busy = 0.0; instr = 0.0; ideal = sum(d[d > 0].sum() for d in tiles) / 32.0
    loading = 0
    ev = [(0.0, w, None) for w in range(W)]
    heapq.heapify(ev)
    live = 0; now = 0.0; idle_time = 0.0; partial_pops = 0; full_pops = 0; dense_pk = 0
    def run(q, pk):
        """pk: list of (slot, death). returns (duration, instr, q_end, survivors list, deaths list)"""
        dur = 0
        while True:
            c0 = 0 if q == 0 else sched[q - 1]; c1 = sched[q]
            dmax = max(min(d, c1) for _, d in pk) - c0
            dur += max(dmax, 0)
            surv = [(s, d) for s, d in pk if d > c1]
            dead = [(s, d) for s, d in pk if d <= c1]
            q += 1
            for s, _ in dead: deaths.append(s)
            pk = surv
            if q == nph or not pk: 
                for s, _ in pk: deaths.append(s)
                return dur, q, []
            if len(pk) >= carry_min: continue
            return dur, q, pk
    while ev:
        now, w, pend = heapq.heappop(ev)
        if pend is not None:
            kind, payload = pend
            if kind == "load":
                slots[payload]["ready"] = True; loading -= 1
            else:
                q, surv, dl = payload
                for s in dl:
                    slots[s]["ref"] -= 1
                    if slots[s]["ref"] == 0: slots[s] = None
                live -= len(dl)
                if surv:
                    if len(buckets[q]) + len(surv) <= cap: buckets[q].extend(surv)
                    else:  # run on immediately
                        deaths = []
                        dur, q2, s2 = run(q, surv)
                        t_busy += dur; instr += dur * NW
                        heapq.heappush(ev, (now + dur, w, ("run", (q2, s2, deaths)))); continue
        # pick
        full = [q for q in range(nph) if len(buckets[q]) >= G]
        pk = None; q = None
        if full:
            q = max(full); pk = buckets[q][-G:]; del buckets[q][-G:]; full_pops += 1
        else:
            free = [i for i in range(R) if slots[i] is None]
            if free and next_tile < len(tiles):
                d = tiles[next_tile]; next_tile += 1
                nv = int((d > 0).sum())
                slots[free[0]] = dict(d=d, cursor=0, npk=(len(d) + G - 1) // G, ref=nv, ready=False)
                live += nv; loading += 1
                heapq.heappush(ev, (now + load_time, w, ("load", free[0]))); continue
            for i in range(R):
                s = slots[i]
                if s is not None and s["ready"] and s["cursor"] < s["npk"]:
                    c = s["cursor"]; s["cursor"] += 1
                    dd = s["d"][c * G:(c + 1) * G]
                    pk = [(i, int(x)) for x in dd if x > 0]; q = 0; dense_pk += 1
                    if not pk: pk = None; continue
                    break
            if pk is None:
                ne = [q for q in range(nph) if buckets[q]]
                if ne:
                    q = max(ne); pk = buckets[q][-G:]; del buckets[q][-G:]; partial_pops += 1
        if pk is None:
            if next_tile >= len(tiles) and live == 0 and loading == 0: continue   # warp done
            idle_time += 0.5
            heapq.heappush(ev, (now + 0.5, w, None)); continue
        deaths = []
        dur, q2, s2 = run(q, pk)
        t_busy += dur; instr += dur * NW
        heapq.heappush(ev, (now + dur, w, ("run", (q2, s2, deaths))))
    return dict(makespan=now, util=t_busy / (W * now), instr_vs_ideal=instr / ideal, ideal=ideal,
                thr=ideal / now, full=full_pops, partial=partial_pops, dense=dense_pk)

if __name__ == "__main__":
    o = pyoracle.Oracle(); h = o.load("tests/golden/jda_shipped_f32.model", False)
    for name, mk in [("noise", synth.noise_frame), ("blur6", synth.blur_frame)]:
        T = []
        for s in range(3):
            tn, _, _ = o.trace(h, mk(s), max_size=192, t_limit=1)
            T += [d for smem, win, d in tiles_of_frame(tn, 640, 480, 192) if win == 24]
        print(name, len(T), "tiles of win 24")
        for W, R, NW in [(24, 12, 2), (24, 12, 1), (16, 12, 2), (24, 6, 2), (24, 24, 2)]:
            r = simulate(T, W=W, R=R, NW=NW)
            print("  W=%d R=%d NW=%d: instr %.2fx ideal, warp util %.2f, packet-carts/time %.1f (full %d partial %d dense %d)" %
                  (W, R, NW, r["instr_vs_ideal"], r["util"], r["thr"], r["full"], r["partial"], r["dense"]))
        # v1-like reference: independent warps, one tile each, NW=2, time = iterations
        from tests.design_sims.sim_lanes import cost_current
        c, i = cost_current([(True, 24, d) for d in T], 2)
        print("  v1 (12 independent warps): instr %.2fx ideal, packet-carts/time %.1f" % (c / i, i / (c / 2 / 12)))
