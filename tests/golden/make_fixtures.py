"""Regenerates tests/golden/* from the reference.  Runs ONLY in the authoring container
(needs /root/reference and oracle/_ref/libjda_ref.so); the outputs are committed because
/root/reference does not exist on the GPU box.

  jda_shipped_f32.model   the shipped model (model/jda_tmp_..._stage_5_cart_540.model) re-serialised
                          by the REFERENCE's own jdaCascadorSerializeTo (c/jda.c:644-716): float32
                          flavour, 5,389,448 B, loads to the bit-identical in-memory cascador.
  face_222x216.npy,       gray crops of model/jda-27.png, cv2 INTER_AREA, the two sizes of the
  face_111x108.npy        SURVEY.md section-4 known-answer canvas.
  ref_outputs.npz         outputs of the REFERENCE jdaDetect on seeded frames (see CASES): the
                          golden vectors every implementation (oracle, CUDA) is checked against.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SHIPPED = os.path.join(REF, "model", "jda_tmp_20160913-112648_stage_5_cart_540.model")

# (name, frame builder, detect kwargs)
CASES = [
    ("noise0", lambda s: s.noise_frame(0), dict(scale=1.25, min_size=24, max_size=-1, th=0.0)),
    ("noise0_3oct", lambda s: s.noise_frame(0), dict(scale=1.25, min_size=24, max_size=192, th=0.0)),
    ("blur1", lambda s: s.blur_frame(1), dict(scale=1.25, min_size=24, max_size=-1, th=0.0)),
    ("facecanvas", lambda s: s.face_canvas(), dict(scale=1.25, min_size=24, max_size=-1, th=0.0)),
    ("facecanvas_main", lambda s: s.face_canvas(), dict(scale=1.25, min_size=40, max_size=-1, th=-0.5)),
    ("facemix3", lambda s: s.facemix_frame(3), dict(scale=1.25, min_size=24, max_size=192, th=0.0)),
    ("facemix5_s12", lambda s: s.facemix_frame(5), dict(scale=1.2, min_size=30, max_size=300, th=-1.0)),
    ("fddb7", lambda s: s.facemix_frame(7, *s.fddb_shape(7)), dict(scale=1.25, min_size=24, max_size=-1, th=0.0)),
    ("tiny30", lambda s: s.blur_frame(2, 30, 27), dict(scale=1.25, min_size=24, max_size=-1, th=-5.0)),
    ("hd_blur", lambda s: s.facemix_frame(11, 1920, 1080), dict(scale=1.25, min_size=24, max_size=768, th=0.0)),
]


def main():
    import cv2
    from oracle import pyoracle
    from jda_b200 import synth

    pyoracle.build()
    ref = pyoracle.RefLib()
    h = ref.load(SHIPPED, double=True)
    out_model = os.path.join(HERE, "jda_shipped_f32.model")
    ref.save_f32(h, out_model)
    print("wrote", out_model, os.path.getsize(out_model))

    img = cv2.imread(os.path.join(REF, "model", "jda-27.png"), cv2.IMREAD_GRAYSCALE)
    np.save(os.path.join(HERE, "face_222x216.npy"), cv2.resize(img, (222, 216), interpolation=cv2.INTER_AREA))
    np.save(os.path.join(HERE, "face_111x108.npy"), cv2.resize(img, (111, 108), interpolation=cv2.INTER_AREA))

    gold = {}
    for name, mk, kw in CASES:
        frame = mk(synth)
        boxes, scores, shapes = ref.detect(h, frame, **kw)
        gold[name + "/boxes"] = boxes
        gold[name + "/scores"] = scores
        gold[name + "/shapes"] = shapes
        gold[name + "/crc"] = np.array([int(frame.astype(np.uint64).sum()), frame.shape[1], frame.shape[0]])
        print(name, frame.shape, "n =", len(scores), scores[:4], boxes[:4].tolist())
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **gold)
    ref.release(h)


if __name__ == "__main__":
    main()
