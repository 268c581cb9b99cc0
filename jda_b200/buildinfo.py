"""Identity of the kernel sources, used to tie an ncu capture (profiles/k2_capture.json) to the tree a bench
line was produced from: the GPU box gets a snapshot without .git, so the tie is a hash of csrc/ itself."""
import hashlib
import os

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def source_sha():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".hpp", ".h")):
            h.update(name.encode())
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]
