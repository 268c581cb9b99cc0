"""k2_scan time per pyramid level: one level at a time (min_size = max_size = win) on 256 resident mix frames."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jda_b200 import api, synth
# one level of 256 frames can be under the small-batch thresholds: the probe is about the throughput tile plan
os.environ.setdefault("JDA_B200_FORCE_PLAN", "throughput")
c = api.Cascador("tests/golden/jda_shipped_f32.model", double=False)
pool = synth.make_frames("mix", 64, 640, 480, seed0=100000)
fr = np.ascontiguousarray(pool[np.arange(256) % 64])
d = torch.from_numpy(fr).cuda()
torch.cuda.synchronize()
wins = api.levels(640, 480, 1.25, 24, 192)
plan = {p["win"]: p for p in api.describe_plan(640, 480, 1.25, 24, 192)}
tot = 0.0
once = "--once" in sys.argv  # one launch per level (+ one of all levels): the launch list ncu sees is [levels..., all]
for w in wins:
    kw = dict(scale=1.25, min_size=w, max_size=w, th=0.0)
    for _ in range(0 if once else 2):
        c.detect_batch(None, device_ptr=d.data_ptr(), shape=(256, 480, 640), unpack=False, **kw)
    ms = []
    for _ in range(1 if once else 4):
        c.detect_batch(None, device_ptr=d.data_ptr(), shape=(256, 480, 640), unpack=False, **kw)
        ms.append(c.last_stats["ms_scan"])
    st = c.last_stats
    p = plan[w]
    k2 = float(np.median(ms)); tot += k2
    print("win %3d step %2d: %7d windows/frame  k2 %.3f ms  %.2f Gwin/s  survivors %d  [%s tile %dx%d span %d]" %
          (w, p["step"], st["windows"] // 256, k2, st["windows"] / k2 / 1e6, st["stage0_survivors"],
           "smem" if p["smem"] else "global", p["tw"], p["th"], p["span"]))
c.detect_batch(None, device_ptr=d.data_ptr(), shape=(256, 480, 640), unpack=False, scale=1.25, min_size=24, max_size=192, th=0.0)
print("sum of single-level scans %.3f ms; all levels in one launch %.3f ms" % (tot, c.last_stats["ms_scan"]))
