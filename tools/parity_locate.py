"""Locate the worst channel differences of one seeded frame (GPU box): scale, channel, position, values."""
import sys

import numpy as np

import acf_b200
from acf_b200 import synth
from oracle.oracle import Oracle


def main():
    seed = int(sys.argv[1]); rows = int(sys.argv[2]); cols = int(sys.argv[3])
    opts = synth.face_opts(80)
    orc = Oracle("port")
    clf = synth.make_classifier(opts, 16, 2, seed=1)
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=1)
    img = synth.shapes_frame(seed, rows, cols)
    taps = {}
    Po = orc.pyramid(opts, img, taps=taps)
    Pg = det.computePyramid(img)
    for i, (g, o) in enumerate(zip(Pg.data, Po.data)):
        d = np.abs(g - o)
        if d.max() > 1e-4:
            idx = np.argwhere(d > 1e-4)
            print("scale", i, "shape", g.shape, "n>1e-4:", len(idx), "max", d.max())
            for z, x, y in idx[:6]:
                print("   chn", z, "x", x, "y", y, "gpu", g[z, x, y], "oracle", o[z, x, y])
    # real-scale taps: smoothed image and real channels of octave 0
    C = det.tap("C", 0, 0, (1, cols, rows))
    Co = taps[("C", 0)]
    dc = np.abs(C - Co)
    print("C0 max diff", dc.max(), "n != :", int((C != Co).sum()), "rows of differing pixels (mod 96):", sorted(set((np.argwhere(C != Co)[:, 2] % 96).tolist()))[:20])
    R = det.tap("R", 0, 0, (7, cols // 4, rows // 4))
    H = taps[("H", 0)]
    dh = np.abs(R[1:] - H)
    print("H0 max diff", dh.max(), "at", np.unravel_index(dh.argmax(), dh.shape))


if __name__ == "__main__":
    main()
