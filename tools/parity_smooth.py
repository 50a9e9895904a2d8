"""Check the engine's smoothed image C against the recurrence evaluated in numpy float32 from the engine's own I and C."""
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
    orc.pyramid(opts, img, taps=taps)
    det.computePyramid(img)
    I = det.tap("I", 0, 0, (1, cols, rows))[0]   # [x, y]
    C = det.tap("C", 0, 0, (1, cols, rows))[0]
    Io = taps[("I", -1)][0]; Co = taps[("C", 0)][0]
    print("I equal:", np.array_equal(I, Io), " C differing:", int((C != Co).sum()))
    f = np.float32
    p, nrm = f(2), f(1 / 16)
    # emulate column by column with the ORACLE's previous column (to isolate the first divergence per column)
    bad = np.argwhere(C != Co)
    xs = np.unique(bad[:, 0])
    print("first differing columns:", xs[:10], "of", len(xs))
    x = int(xs[0])
    ys = bad[bad[:, 0] == x][:, 1]
    print("column", x, "differing rows:", ys[:20])
    prev = Co[x - 1] if x > 0 else Io[x]
    cur = Io[x]; nxt = Io[x + 1] if x + 1 < cols else Io[x]
    T = nrm * ((prev + p * cur) + nxt)
    O = np.empty_like(T)
    O[1:-1] = (T[:-2] + p * T[1:-1]) + T[2:]
    O[0] = (f(1) + p) * T[0] + T[1]; O[-1] = T[-2] + (f(1) + p) * T[-1]
    print("numpy emulation == oracle column:", np.array_equal(O, Co[x]), " == gpu column:", np.array_equal(O, C[x]))
    y = int(ys[0])
    print("row", y, "gpu", repr(C[x, y]), "oracle", repr(Co[x, y]), "T around:", T[y - 1:y + 2], "prev/cur/nxt:", prev[y], cur[y], nxt[y])
    print("gpu prev col equal oracle prev col:", np.array_equal(C[x - 1], Co[x - 1]) if x > 0 else None)


if __name__ == "__main__":
    main()


def locate_hist():
    seed = int(sys.argv[1]); rows = int(sys.argv[2]); cols = int(sys.argv[3])
    opts = synth.face_opts(80)
    orc = Oracle("port")
    clf = synth.make_classifier(opts, 16, 2, seed=1)
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=1)
    img = synth.shapes_frame(seed, rows, cols)
    taps = {}
    orc.pyramid(opts, img, taps=taps)
    det.computePyramid(img)
    C = det.tap("C", 0, 0, (1, cols, rows))[0]
    Co = taps[("C", 0)][0]
    R = det.tap("R", 0, 0, (7, cols // 4, rows // 4))
    H = taps[("H", 0)]
    dh = np.abs(R[1:] - H)
    b, cx, cy = np.unravel_index(dh.argmax(), dh.shape)
    print("worst H cell: bin", b, "cell x", cx, "cell y", cy, "gpu", R[1 + b, cx, cy], "oracle", H[b, cx, cy])
    print("all bins gpu   ", R[1:, cx, cy]); print("all bins oracle", H[:, cx, cy])
    x0, y0 = 4 * cx, 4 * cy
    sl = (slice(max(0, x0 - 1), x0 + 5), slice(max(0, y0 - 1), y0 + 5))
    print("C gpu == oracle in the neighbourhood:", np.array_equal(C[sl], Co[sl]), " n differing", int((C[sl] != Co[sl]).sum()))
    O = taps[("O", 0)][0]; M = taps[("M", 0)][0] if ("M", 0) in taps else None
    print("oracle O in cell:\n", O[x0:x0 + 4, y0:y0 + 4])
    if M is not None: print("oracle M in cell:\n", M[x0:x0 + 4, y0:y0 + 4])
    print("oracle C:\n", Co[sl]); print("gpu C - oracle C:\n", C[sl] - Co[sl])
    print("rows mod 96:", y0 % 96, " x:", x0, "of", cols)


if len(sys.argv) > 4:
    locate_hist()


def self_consistency():
    seed = int(sys.argv[1]); rows = int(sys.argv[2]); cols = int(sys.argv[3])
    opts = synth.face_opts(80)
    clf = synth.make_classifier(opts, 16, 2, seed=1)
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=1)
    img = synth.shapes_frame(seed, rows, cols)
    det.computePyramid(img)
    I = det.tap("I", 0, 0, (1, cols, rows))[0]
    C = det.tap("C", 0, 0, (1, cols, rows))[0]
    f = np.float32
    p, nrm = f(2), f(1 / 16)
    tot = 0; big = 0; worst = 0.0; where = None
    import collections
    hist = collections.Counter(); cols_bad = collections.Counter()
    for x in range(cols):
        prev = C[x - 1] if x > 0 else I[x]
        cur = I[x]; nxt = I[x + 1] if x + 1 < cols else I[x]
        T = nrm * ((prev + p * cur) + nxt)
        O = np.empty_like(T)
        O[1:-1] = (T[:-2] + p * T[1:-1]) + T[2:]
        O[0] = (f(1) + p) * T[0] + T[1]; O[-1] = T[-2] + (f(1) + p) * T[-1]
        bad = np.nonzero(O != C[x])[0]
        tot += len(bad)
        if len(bad):
            d = np.abs(O[bad] - C[x][bad])
            big += int((d > 1e-9).sum())
            for r in bad[d > 1e-9]: hist[int(r) % 96] += 1
            if (d > 1e-9).any(): cols_bad[x] += 1
            if d.max() > worst:
                worst = float(d.max()); where = (x, int(bad[d.argmax()]), float(O[bad[d.argmax()]]), float(C[x][bad[d.argmax()]]))
    print('rows mod 96 of the large mismatches:', sorted(hist.items()))
    print('columns with large mismatches (first 30):', sorted(cols_bad)[:30], 'n', len(cols_bad))
    print("pixels where the GPU column differs from the formula applied to its own previous column:", tot, " with |d|>1e-9:", big, "worst", worst, where)


if len(sys.argv) > 5:
    self_consistency()


def vs_oracle_hist():
    import collections
    seed = int(sys.argv[1]); rows = int(sys.argv[2]); cols = int(sys.argv[3])
    opts = synth.face_opts(80)
    orc = Oracle("port")
    clf = synth.make_classifier(opts, 16, 2, seed=1)
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=1)
    img = synth.shapes_frame(seed, rows, cols)
    taps = {}
    orc.pyramid(opts, img, taps=taps)
    det.computePyramid(img)
    C = det.tap("C", 0, 0, (1, cols, rows))[0]
    Co = taps[("C", 0)][0]
    bad = np.argwhere(np.abs(C - Co) > 1e-9)
    print("n |d|>1e-9:", len(bad))
    h = collections.Counter((bad[:, 1] % 96).tolist())
    print("by row mod 96:", sorted(h.items()))
    # first column where each differing row starts to differ
    first = {}
    for x, y in bad:
        first.setdefault(int(y), int(x))
    fr = sorted(first.items(), key=lambda kv: kv[1])[:12]
    print("rows by first differing column:", fr)
    for y, x in fr[:3]:
        print(" row", y, "col", x, "gpu", C[x - 1:x + 2, y], "oracle", Co[x - 1:x + 2, y], "rows above/below oracle", Co[x, y - 2:y + 3])


if len(sys.argv) > 6:
    vs_oracle_hist()
