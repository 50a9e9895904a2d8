"""GPU results against the reference AS SHIPPED (protocol step 3).

The engine is bit-identical to the reference's exact-math build (rcpps / rsqrtps replaced by IEEE 1/x, 1/sqrt; tests/).  The
shipped build uses the SSE approximations, so its channels differ by up to ~5e-4 and a window whose deciding
feature lies that close to a threshold can fall the other way.  This tool runs the same frames through the GPU and through
oracle/_ref/liboracle_ref_native.so (the reference's own toolbox objects, unmodified) and reports, per frame: raw hits of both,
the symmetric difference of the hit sets (scale, c, r), and the largest score difference over common hits.
Usage (GPU box): PYTHONPATH=. python tools/native_deviation.py [n_frames] [out.md]"""
import sys

import numpy as np

import acf_b200
from acf_b200 import synth
from oracle import oracle as O


def compare(det, orc, opts, clf, frame):
    det(frame, cap=1 << 20)
    hits, trees, windows = det.last_hits()
    g = {(h[1], h[2], h[3]): h[4] for h in hits}
    P = orc.pyramid(opts, frame)
    dets, (hs, hc, hr), ne, total = P.detect(clf, cap=1 << 20)
    chan = max(float(np.abs(a - b).max()) for a, b in zip(det.readPyramid(frame.shape[0], frame.shape[1]).data, P.data))
    P.close()
    o = {(int(a), int(b), int(c)): np.float32(d[4]) for a, b, c, d in zip(hs, hc, hr, dets)}
    common = set(g) & set(o)
    dmax = max((abs(float(g[k]) - float(o[k])) for k in common), default=0.0)
    return dict(gpu=len(g), native=len(o), only_gpu=len(set(g) - set(o)), only_native=len(set(o) - set(g)), max_score_delta=dmax,
                trees_gpu=trees, trees_native=ne, windows=windows, max_channel_delta=chan)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    out = sys.argv[2] if len(sys.argv) > 2 else None
    if not O.available("ref_native"):
        raise SystemExit("oracle/_ref/liboracle_ref_native.so is not built (needs /root/reference at build time)")
    orc = O.Oracle("ref_native")
    opts = synth.face_opts(80)
    clf = synth.make_classifier(opts, 2048, 2, seed=1, n_reject=54)  # bench.py's headline model
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=1080, max_cols=1920, max_batch=1)
    det.setHitCapacity(1 << 17)
    rows = []
    for i in range(n):
        r = compare(det, orc, opts, clf, synth.shapes_frame(100 + i, 1080, 1920))
        rows.append(r)
        print(i, r, flush=True)
    tot = {k: sum(r[k] for r in rows) for k in ("gpu", "native", "only_gpu", "only_native", "windows")}
    lines = ["# GPU (= reference exact-math build) vs the reference as shipped (SSE rcpps / rsqrtps)", "",
             "`tools/native_deviation.py`: bench.py's headline workload (1080p 'shapes' frames 100.., FACE80, 2048 trees, n_reject 54), "
             "one frame at a time; hit = (scale, c, r).", "",
             "| frame | raw hits GPU | raw hits shipped build | only GPU | only shipped | max score delta (common hits) | max channel delta | trees GPU | trees shipped |",
             "|---|---|---|---|---|---|---|---|---|"]
    for i, r in enumerate(rows):
        lines.append(f"| {100 + i} | {r['gpu']} | {r['native']} | {r['only_gpu']} | {r['only_native']} | {r['max_score_delta']:.3e} | {r['max_channel_delta']:.3e} | {r['trees_gpu']} | {r['trees_native']} |")
    lines += ["", f"Totals over {n} frames ({tot['windows']} windows): {tot['gpu']} hits on the GPU, {tot['native']} in the shipped build; "
              f"{tot['only_gpu']} only on the GPU, {tot['only_native']} only in the shipped build "
              f"({100.0 * (tot['only_gpu'] + tot['only_native']) / max(1, tot['gpu'] + tot['native']):.2f} % of all hits, "
              f"{1e6 * (tot['only_gpu'] + tot['only_native']) / max(1, tot['windows']):.1f} per million windows)."]
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
