"""Generates tests/golden/golden_v1.npz from the REFERENCE's own toolbox objects
(oracle/_ref/liboracle_ref_exact.so, built by `make -C oracle ref` from /root/reference) under the
restated orchestration.  Run in the build container (needs /root/reference); the fixture travels.

Contents: per-stage outputs of the reference's L1 functions on a seeded 48x40 image, resample
cases, a full small pyramid (gray 7-channel and LUV 10-channel padded) and the detection list of a
seeded synthetic classifier.  The reference's own tests hold no numeric vectors for this path
(SURVEY.md 4), so these outputs of the reference code are the pin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from acf_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def small_face_opts():
    o = synth.face_opts(32)
    return o


def small_inria_opts():
    o = synth.inria_opts()
    o.update(minDs=(48, 24), modelDs=(48, 24), modelDsPad=(64, 32), pad=(8, 4))
    return o


def main():
    E = Oracle("ref_exact")
    N = Oracle("ref_native")
    rng = np.random.default_rng(1234)
    out = {}
    I = rng.random((3, 40, 48), dtype=np.float32)  # [plane, w, h]
    I[:, :6, :6] = 0.25
    out["l1_rgb"] = I
    out["l1_gray"] = E.rgb_convert(I, 0)
    out["l1_luv"] = E.rgb_convert(I, 2)
    out["l1_luv_native"] = N.rgb_convert(I, 2)
    g = out["l1_gray"]
    out["l1_tri1_oop"] = E.conv_tri1(g, 2.0)
    out["l1_tri1_inplace"] = E.conv_tri1(g, 2.0, True)
    out["l1_tri5"] = E.conv_tri(g, 5)
    C = out["l1_tri1_inplace"][0]
    for full in (0, 1):
        M, O = E.grad_mag(C, full)
        out[f"l1_M_full{full}"] = M; out[f"l1_O_full{full}"] = O
    M, O = E.grad_mag(C, 0)
    Mn_native, _ = N.grad_mag(C, 0)
    out["l1_M_native"] = Mn_native
    S = E.conv_tri(M[None], 5)[0]
    Mn = E.grad_mag_norm(M, S, 0.005)
    out["l1_S"] = S; out["l1_Mnorm"] = Mn
    out["l1_H"] = E.grad_hist(Mn, O, 4, 6, 0, 0)
    out["l1_H_hard"] = E.grad_hist(Mn, O, 4, 6, -2, 0)
    A = rng.random((2, 60, 48), dtype=np.float32)
    out["rs_src"] = A
    for (wb, hb) in [(30, 24), (15, 12), (20, 16), (43, 34), (16, 13), (75, 60), (60, 48), (11, 9)]:
        out[f"rs_{wb}x{hb}"] = E.resample(A, hb, wb, 1.3)
    # full pyramids + detections
    for name, opts, frame in (("face", small_face_opts(), synth.noise_frame(7, 128, 160)),
                              ("inria", small_inria_opts(), synth.shapes_frame(11, 160, 128))):
        P = E.pyramid(opts, frame)
        out[f"{name}_frame"] = frame
        out[f"{name}_scales"] = np.array(P.scales)
        out[f"{name}_scaleshw"] = np.array(P.scaleshw)
        for i, d in enumerate(P.data):
            out[f"{name}_pyr{i:02d}"] = d
        clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
        dets, (hs, hc, hr), ne, total = P.detect(clf)
        out[f"{name}_dets"] = np.array([[d[0], d[1], d[2], d[3]] for d in dets], np.int32).reshape(-1, 4)
        out[f"{name}_scores"] = np.array([d[4] for d in dets], np.float64)
        out[f"{name}_hits"] = np.stack([hs, hc, hr], 1).astype(np.int32)
        out[f"{name}_trees"] = np.array([ne], np.int64)
        print(name, "scales", P.nScales, "dets", total, "trees", ne)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
