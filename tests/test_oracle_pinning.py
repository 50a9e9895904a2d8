"""CPU: pins the oracle.  (1) the port restatement against golden vectors produced by the
REFERENCE's own toolbox objects (tests/golden/make_golden.py); (2) when oracle/_ref is present,
port == reference objects bit for bit on fresh seeded inputs; (3) native (rcpps/rsqrtps) vs exact
deviation stays inside the band SURVEY.md 0.6 measured."""
import numpy as np
import pytest

from acf_b200 import synth
from tests.golden.make_golden import small_face_opts, small_inria_opts


def test_port_l1_matches_reference_golden(oracle_port, golden):
    P, G = oracle_port, golden
    I = G["l1_rgb"]
    assert np.array_equal(P.rgb_convert(I, 0), G["l1_gray"])
    assert np.array_equal(P.rgb_convert(I, 2), G["l1_luv"])
    g = G["l1_gray"]
    assert np.array_equal(P.conv_tri1(g, 2.0), G["l1_tri1_oop"])
    assert np.array_equal(P.conv_tri1(g, 2.0, True), G["l1_tri1_inplace"])
    assert np.array_equal(P.conv_tri(g, 5), G["l1_tri5"])
    C = G["l1_tri1_inplace"][0]
    for full in (0, 1):
        M, O = P.grad_mag(C, full)
        assert np.array_equal(M, G[f"l1_M_full{full}"]) and np.array_equal(O, G[f"l1_O_full{full}"])
    M, O = P.grad_mag(C, 0)
    Mn = P.grad_mag_norm(M, G["l1_S"], 0.005)
    assert np.array_equal(Mn, G["l1_Mnorm"])
    assert np.array_equal(P.grad_hist(Mn, O, 4, 6, 0, 0), G["l1_H"])
    assert np.array_equal(P.grad_hist(Mn, O, 4, 6, -2, 0), G["l1_H_hard"])
    A = G["rs_src"]
    for key in [k for k in G.files if k.startswith("rs_") and k != "rs_src"]:
        wb, hb = map(int, key[3:].split("x"))
        assert np.array_equal(P.resample(A, hb, wb, 1.3), G[key]), key


def test_in_place_smoothing_is_a_recurrence(oracle_port, golden):
    # aliased convTri1 differs from the out-of-place filter (probe: up to 0.08)
    d = np.abs(golden["l1_tri1_inplace"] - golden["l1_tri1_oop"]).max()
    assert 1e-3 < d < 0.2


@pytest.mark.parametrize("name,opts_fn", [("face", small_face_opts), ("inria", small_inria_opts)])
def test_port_pyramid_and_detections_match_reference_golden(oracle_port, golden, name, opts_fn):
    opts = opts_fn()
    P = oracle_port.pyramid(opts, golden[f"{name}_frame"])
    assert np.array_equal(np.array(P.scales), golden[f"{name}_scales"])
    assert np.array_equal(np.array(P.scaleshw), golden[f"{name}_scaleshw"])
    for i, d in enumerate(P.data):
        assert np.array_equal(d, golden[f"{name}_pyr{i:02d}"]), f"scale {i}"
    clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    dets, (hs, hc, hr), ne, total = P.detect(clf)
    assert total == len(golden[f"{name}_dets"]) and total > 0
    assert np.array_equal(np.array([d[:4] for d in dets], np.int32).reshape(-1, 4), golden[f"{name}_dets"])
    assert np.array_equal(np.array([d[4] for d in dets]), golden[f"{name}_scores"])
    assert np.array_equal(np.stack([hs, hc, hr], 1), golden[f"{name}_hits"])
    assert ne == int(golden[f"{name}_trees"][0])


def test_port_equals_reference_objects_on_fresh_inputs(oracle_port, oracle_ref_exact):
    rng = np.random.default_rng(99)
    P, E = oracle_port, oracle_ref_exact
    for shape in [(3, 52, 36), (3, 33, 47)]:
        I = rng.random(shape, dtype=np.float32)
        g = P.rgb_convert(I, 0)
        assert np.array_equal(g, E.rgb_convert(I, 0))
        if (shape[1] * shape[2]) % 4 == 0:  # the reference takes its SSE luv path only when n % 4 == 0
            assert np.array_equal(P.rgb_convert(I, 2), E.rgb_convert(I, 2))
        # hsv (rgbConvertMex.cpp:193-238) incl. gray pixels and ties between the channel maxima
        Q = (I * 8).astype(np.int32).astype(np.float32) / 8
        assert np.array_equal(P.rgb_convert(Q, 3), E.rgb_convert(Q, 3))
        assert np.array_equal(P.rgb_convert(I, 3), E.rgb_convert(I, 3))
        assert np.array_equal(P.conv_tri1(g, 2.0, True), E.conv_tri1(g, 2.0, True))
        assert np.array_equal(P.conv_tri(g, 5), E.conv_tri(g, 5))
        Mp, Op = P.grad_mag(g[0], 0); Me, Oe = E.grad_mag(g[0], 0)
        assert np.array_equal(Mp, Me) and np.array_equal(Op, Oe)
    for rows, cols, opts in [(96, 128, small_face_opts()), (128, 96, small_inria_opts()),
                             (100, 132, dict(small_face_opts(), lambdas=[])),
                             (96, 128, dict(small_inria_opts(), colorSpace="hsv", gm_colorChn=2)),
                             (96, 128, dict(small_inria_opts(), colorSpace="rgb", gm_colorChn=1))]:
        img = synth.noise_frame(3, rows, cols)
        a, b = P.pyramid(opts, img), E.pyramid(opts, img)
        assert a.nScales == b.nScales and a.lambdas == b.lambdas
        for x, y in zip(a.data, b.data):
            assert np.array_equal(x, y)


def test_native_sse_deviation_band(oracle_ref_exact, oracle_ref_native, golden):
    # rcpps / rsqrtps (toolbox/sse.hpp:185-192) vs IEEE: M up to ~5.4e-4 relative
    C = golden["l1_tri1_inplace"][0]
    Me, _ = oracle_ref_exact.grad_mag(C, 0)
    Mn, _ = oracle_ref_native.grad_mag(C, 0)
    rel = np.abs(Mn - Me) / np.maximum(Me, 1e-6)
    assert 1e-6 < rel.max() < 2e-3
    assert np.abs(golden["l1_luv_native"] - golden["l1_luv"]).max() < 2e-3


def test_get_scales_known_values(oracle_port):
    # SURVEY Appendix C probe: real scales for 1080p / 640x480; s == 0.5 compares true
    s, hw = oracle_port.get_scales(8, 0, (80, 80), 4, (1080, 1920))
    assert len(s) == 31 and s[0] == 1.0 and s[8] == 0.5
    assert abs(s[16] - 0.252) < 5e-4 and abs(s[24] - 0.12533) < 5e-4
    s, hw = oracle_port.get_scales(8, 0, (64, 64), 4, (640, 480))
    assert len(s) == 24 and s[8] == 0.5 and s[16] == 0.25


def test_nms_and_prune(oracle_port):
    dets = [(10, 10, 40, 40, 5.0), (12, 12, 40, 40, 4.0), (100, 100, 40, 40, 3.0), (14, 10, 40, 40, 6.0)]
    kept = oracle_port.nms(dets, overlap=0.5, greedy=True, ovr_union=True)
    assert [k[4] for k in kept] == [6.0, 3.0]
    # prune keeps boxes until one falls below ratio * best score (that one is still kept, ObjectDetector.cpp:33-40)
    assert len(oracle_port.prune(kept + [(0, 0, 1, 1, 0.1), (5, 5, 1, 1, 0.05)], max_count=10, ratio=0.4)) == 3
    assert len(oracle_port.prune(kept, max_count=1, ratio=0.0)) == 1
