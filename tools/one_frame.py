import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jda_b200 import api, synth
c = api.Cascador("tests/golden/jda_shipped_f32.model", double=False)
img = synth.face_canvas() if len(sys.argv) < 2 or sys.argv[1] == "faces" else synth.noise_frame(1)
for _ in range(6):
    c.detect(img, 1.25, 0.1, 24, 192, 0.0)
