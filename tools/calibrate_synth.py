"""One-off helper: empirical quantiles of pyramid channel values on the synthetic frames, and
mean trees/window of the synthetic classifier, measured through the CPU oracle.  Used to fill
acf_b200/synth.py:_QUANT and to choose make_classifier's defaults.  Not part of the product."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acf_b200 import synth
from oracle.oracle import Oracle

def main():
    orc = Oracle("port")
    rows, cols = 480, 640
    for name, opts in (("face8", synth.face_opts(64, True)), ("inria", synth.inria_opts())):
        vals = {}
        for kind in ("shapes", "noise"):
            for seed in range(3):
                img = getattr(synth, kind + "_frame")(seed, rows, cols)
                P = orc.pyramid(opts, img)
                kinds = synth.channel_kinds(opts)
                for d in P.data[:12]:
                    for z, k in enumerate(kinds):
                        vals.setdefault(k, []).append(d[z].ravel())
        for k, v in vals.items():
            v = np.concatenate(v)
            print(name, k, np.round(np.percentile(v, [5, 25, 50, 75, 95]), 4))

if __name__ == "__main__":
    main()
