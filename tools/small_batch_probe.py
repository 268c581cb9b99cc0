"""Tile plan and cascade kernels on small batches (what coalesced jdaDetect calls and small jdaB200DetectBatch calls run):
stage kernels (k3_walk / k3_regress / k3_emit) against k3_cascade for stages >= 1, resident frames, steady state."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jda_b200 import api, synth


def handle(env=None):
    if env:
        os.environ[env[0]] = env[1]
    try:
        return api.Cascador("tests/golden/jda_shipped_f32.model", double=False)
    finally:
        if env:
            del os.environ[env[0]]


pool = synth.make_frames("mix", 64, 640, 480, seed0=100000)
for name, env in (("default", None), ("stage kernels", ("JDA_B200_STAGE_MIN_WINDOWS", "0")),
                  ("k3_cascade for stages >= 1", ("JDA_B200_NO_STAGE_KERNELS", "1")),
                  ("latency tile plan", ("JDA_B200_FORCE_PLAN", "latency"))):
    c = handle(env)
    for n in ((1, 2, 4, 5, 8, 12, 16, 24, 32, 64) if name in ("default", "latency tile plan") else (5, 16, 64, 256)):
        fr = np.ascontiguousarray(pool[np.arange(n) % 64])
        d = torch.from_numpy(fr).cuda()
        torch.cuda.synchronize()
        kw = dict(scale=1.25, min_size=24, max_size=192, th=0.0)
        for _ in range(3):
            c.detect_batch(None, device_ptr=d.data_ptr(), shape=(n, 480, 640), unpack=False, **kw)
        sc, ca = [], []
        t0 = time.perf_counter()
        for _ in range(10):
            c.detect_batch(None, device_ptr=d.data_ptr(), shape=(n, 480, 640), unpack=False, **kw)
            sc.append(c.last_stats["ms_scan"]); ca.append(c.last_stats["ms_cascade"])
        wall = (time.perf_counter() - t0) / 10 * 1e3
        print("%-28s %4d frames: scan %.3f ms  cascade %.3f ms  call %.3f ms  (%d launches)" %
              (name, n, np.median(sc), np.median(ca), wall, c.last_stats["cascade_launches"]), flush=True)
    c.close()
