"""A/B check of two builds of libacf_b200.so: SHA-256 of every pyramid scale for a few seeded frames.

Usage (GPU box):  python tools/ab_pyramid.py [path/to/other/libacf_b200.so]
Prints one digest per (config, scale).  Run it with two libraries and diff the outputs to prove a kernel
rewrite is bit-identical (used when k_chan was specialised).
"""
import hashlib
import sys

import numpy as np

import acf_b200._capi as _capi

if len(sys.argv) > 1:
    _capi.LIB_PATH = sys.argv[1]
from acf_b200 import Detector, Model, synth  # noqa: E402


def main():
    cases = [("face", 540, 960), ("face", 1080, 1920), ("inria", 480, 640), ("face", 250, 333), ("inria", 1080, 1920)]
    for kind, rows, cols in cases:
        opts = synth.face_opts() if kind == "face" else synth.inria_opts()
        clf = synth.make_classifier(opts, n_trees=64, seed=3)
        det = Detector(Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=2)
        frames = np.stack([synth.shapes_frame(11, rows, cols), synth.noise_frame(12, rows, cols)])
        for f in range(2):
            P = det.computePyramid(frames, frame=f)
            for i, d in enumerate(P.data):
                print(kind, rows, cols, f, i, d.shape, hashlib.sha256(d.tobytes()).hexdigest()[:16])
        det.close()


if __name__ == "__main__":
    main()
