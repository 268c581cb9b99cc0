"""single-frame latency breakdown (run on the GPU box)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jda_b200 import api, synth
c = api.Cascador("tests/golden/jda_shipped_f32.model", double=False)
for name, img, mx in [("vga_noise", synth.noise_frame(1), 192), ("vga_faces", synth.face_canvas(), 192),
                      ("hd_facemix", synth.facemix_frame(7000, 1920, 1080), 768)]:
    for _ in range(5):
        c.detect_batch(img[None], max_size=mx)
    ws, ks = [], []
    for _ in range(20):
        t0 = time.perf_counter(); c.detect_batch(img[None], max_size=mx); ws.append((time.perf_counter() - t0) * 1e3)
        ks.append(dict(c.last_stats))
    wd = []
    for _ in range(20):
        t0 = time.perf_counter(); c.detect(img, 1.25, 0.1, 24, mx, 0.0); wd.append((time.perf_counter() - t0) * 1e3)
    k = {key: float(np.median([s[key] for s in ks])) for key in ("ms_h2d", "ms_resize", "ms_scan", "ms_cascade", "ms_d2h", "ms_host")}
    print(name, "wall(batch,stats) p50 %.3f ms | jdaDetect p50 %.3f ms |" % (np.median(ws), np.median(wd)), {a: round(b, 3) for a, b in k.items()},
          "surv", ks[0]["stage0_survivors"], "hits", ks[0]["raw_hits"])
