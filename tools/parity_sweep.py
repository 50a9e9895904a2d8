"""Worst |gpu - oracle| over all scales of the pyramid for a sweep of seeded frames (GPU box).
Usage: PYTHONPATH=. python tools/parity_sweep.py [n_seeds]
The oracle used is the scalar port (bit-identical to the reference's exact-math objects, tests/test_oracle_pinning.py)."""
import sys

import numpy as np

import acf_b200
from acf_b200 import synth
from oracle.oracle import Oracle


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    orc = Oracle("port")
    cases = [("face80 1080p", synth.face_opts(80), 1080, 1920), ("inria 1080p", synth.inria_opts(), 1080, 1920),
             ("face80 4K", synth.face_opts(80), 2160, 3840), ("face64 480p", synth.face_opts(64), 480, 640)]
    for name, opts, rows, cols in cases:
        clf = synth.make_classifier(opts, 16, 2, seed=1)
        det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=1)
        worst = []
        for seed in range(n):
            for kind in ("shapes", "noise"):
                img = getattr(synth, kind + "_frame")(1000 + seed, rows, cols)
                Pg = det.computePyramid(img)
                Po = orc.pyramid(opts, img)
                w = max(float(np.abs(g - o).max()) for g, o in zip(Pg.data, Po.data))
                worst.append((w, kind, seed))
                Po.close()
        worst.sort(reverse=True)
        print(name, "worst", ["%.2e %s#%d" % w for w in worst[:4]], "median %.2e" % worst[len(worst) // 2][0], flush=True)
        det.close()


if __name__ == "__main__":
    main()
