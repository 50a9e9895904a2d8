"""GPU parity of k_cascade_tile (the TMA-staged shared-memory cascade, acf_b200/csrc/cascade_tile.cu) and of the
engine's batch scheduling around it: bit-exact against the CPU oracle (acfDetect1.cpp:84-138) AND against the
global-gather kernel it replaces on the hot path (k_cascade, ACFB_CASC_TILE=0), on the cases that stress what
is new: survivors deep into the streamed table chunks, hits, strides of two channel pixels, 10-channel windows,
scales smaller than a tile, batches split over lanes with three batches in flight, and batch sizes that change
the frame -> lane mapping between batches in flight."""
import os

import numpy as np
import pytest

import acf_b200
from acf_b200 import synth
from tests.golden.make_golden import small_face_opts, small_inria_opts

pytestmark = pytest.mark.gpu


def _detector(opts, clf, tile, rows=512, cols=640, max_batch=4, cap=1 << 18, export=None, sparse=None, tail_tile=None):
    """tile: k_cascade_tile (True) or the global-gather k_cascade (False); export: survivors per tile and level below which the
    tile hands its windows to k_cascade_tail (0 = everything stays in the tile); sparse: the in-tile lanes-as-trees threshold"""
    env = {"ACFB_CASC_TILE": "1" if tile else "0"}
    if export is not None:
        env["ACFB_CASC_EXPORT"] = str(export)
    if sparse is not None:
        env["ACFB_CASC_SPARSE"] = str(sparse)
    if tail_tile is not None:  # hand-over finished on TMA-staged window footprints (k_cascade_tail_win, default) or with global gathers (k_cascade_tail)
        env["ACFB_CASC_TAIL_WIN"] = "1" if tail_tile else "0"
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        det = acf_b200.Detector(acf_b200.Model.create(opts, clf), max_rows=rows, max_cols=cols, max_batch=max_batch)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    det.setHitCapacity(cap)
    return det


def _deep_clf(opts, n_trees, seed=5):
    # rejectors in the first trees, then small positive confirmations: a good share of the windows walks the whole
    # table (hits), the rest leaves at every depth -- every level and both ring slots of the streamed table see work
    return synth.make_classifier(opts, n_trees, 2, seed=seed, drift=-0.04, gain=0.3, n_reject=n_trees // 2, confirm=0.002)


@pytest.mark.parametrize("name,opts_fn,n_trees", [
    ("face32", small_face_opts, 400), ("face32-short", small_face_opts, 40), ("inria", small_inria_opts, 300),
    ("face64-8ch", lambda: synth.face_opts(64, True), 200),
    ("stride8", lambda: dict(small_face_opts(), stride=8), 200),
])
def test_tile_cascade_on_oracle_channels(oracle_port, name, opts_fn, n_trees):
    opts = opts_fn()
    clf = _deep_clf(opts, n_trees)
    variants = (("tile + tail on staged window footprints", _detector(opts, clf, True)),                # the hot path's defaults
                ("tile + tail with global gathers", _detector(opts, clf, True, tail_tile=False)),
                ("tile, tail for everything past tree 64", _detector(opts, clf, True, export=1 << 20)),
                ("tile, gather tail for everything past tree 64", _detector(opts, clf, True, export=1 << 20, tail_tile=False)),
                ("tile only, batches", _detector(opts, clf, True, export=0, sparse=0)),
                ("tile only, lanes as trees", _detector(opts, clf, True, export=0, sparse=1 << 20)),
                ("gather", _detector(opts, clf, False)))
    Po = oracle_port.pyramid(opts, synth.shapes_frame(8, 384, 512))
    nh = deep = 0
    for chns in Po.data:
        oc, or_, os_, one = oracle_port.acf_detect1(chns, opts, clf)
        for what, det in variants:
            c, r, s, ne = det.acfDetect1(chns)
            assert np.array_equal(c, oc) and np.array_equal(r, or_), (name, what)
            assert np.array_equal(s, os_), (name, what, "scores must be bit exact (same sequential float adds)")
            assert ne == one, (name, what, "trees evaluated")
        nh += len(oc)
        nwin = max(1, (chns.shape[1] - opts["modelDsPad"][1] // 4 + 1) * (chns.shape[2] - opts["modelDsPad"][0] // 4 + 1))
        deep = max(deep, one / nwin)
    assert nh > 0
    print(f"{name}: {nh} hits, up to {deep:.1f} trees/window")


@pytest.mark.parametrize("overlap,lanes", [(0, 1), (1, 2)])
def test_tile_cascade_end_to_end_1080p_batches_in_flight(oracle_port, overlap, lanes):
    # What bench.py times: 1080p frames, six batches in flight going round the engine's three pipelines, hits > 0, on the engine's default single stream and with the
    # batch split over two lanes x three streams (ACFB_OVERLAP=1 ACFB_LANES=2) -- every frame must equal its single-frame
    # result bit for bit, and sampled frames the oracle's boxes
    opts = synth.face_opts(80)
    clf = synth.make_classifier(opts, 2048, 2, seed=1, n_reject=54)  # the benchmark model: 0-660 raw hits per frame on these frames
    os.environ["ACFB_OVERLAP"], os.environ["ACFB_LANES"] = str(overlap), str(lanes)
    try:
        det = _detector(opts, clf, True, rows=1080, cols=1920, max_batch=8, cap=1 << 16)
    finally:
        del os.environ["ACFB_OVERLAP"], os.environ["ACFB_LANES"]
    frames = synth.frames("shapes", 8, 1080, 1920, seed0=100)
    batches = [np.ascontiguousarray(frames), np.ascontiguousarray(frames[::-1]), np.ascontiguousarray(np.roll(frames, 3, axis=0)),
               np.ascontiguousarray(np.roll(frames, 5, axis=0)), np.ascontiguousarray(np.roll(frames, 1, axis=0)), np.ascontiguousarray(np.roll(frames, 6, axis=0))]
    single = [det(f, cap=1 << 18) for f in frames]
    assert sum(len(r) for r, _ in single) > 0
    for b in batches:
        det.submit(b.ctypes.data, 8, 1080, 1920, False)
    order = [list(range(8)), list(range(7, -1, -1)), [(i - 3) % 8 for i in range(8)], [(i - 5) % 8 for i in range(8)],
             [(i - 1) % 8 for i in range(8)], [(i - 6) % 8 for i in range(8)]]
    for k in range(6):
        res, total = det.collect(8, cap=1 << 18)
        for i in range(8):
            assert res[i] == single[order[k][i]], (k, i)
    for f in (0, 5):
        odets, _, _, ototal = oracle_port.pyramid(opts, frames[f]).detect(clf)
        assert [tuple(d[:4]) for d in odets] == single[f][0]
        assert np.array_equal(np.array([d[4] for d in odets], np.float32), np.array(single[f][1], np.float32))
    print(f"hits per frame: {[len(r) for r, _ in single]}")


def test_changing_batch_size_between_batches_in_flight():
    # with n = 8 then n = 5 (then 8 again) not collected in between, frames move between lanes while the
    # previous batch's channel / cascade kernels may still read R and the pyramid of those frame slots
    opts = small_face_opts()
    clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    os.environ["ACFB_OVERLAP"], os.environ["ACFB_LANES"] = "1", "2"  # the lane split is what makes the mapping change
    try:
        det = _detector(opts, clf, True, rows=512, cols=640, max_batch=8)
    finally:
        del os.environ["ACFB_OVERLAP"], os.environ["ACFB_LANES"]
    fr = synth.frames("shapes", 8, 480, 640, seed0=70)
    single = [det(f, cap=1 << 18) for f in fr]
    for rep in range(3):
        a, b, c = np.ascontiguousarray(fr), np.ascontiguousarray(fr[3:8]), np.ascontiguousarray(fr[::-1])
        det.submit(a.ctypes.data, 8, 480, 640, False)
        det.submit(b.ctypes.data, 5, 480, 640, False)
        det.submit(c.ctypes.data, 8, 480, 640, False)
        ra, _ = det.collect(8, cap=1 << 18)
        rb, _ = det.collect(5, cap=1 << 18)
        rc, _ = det.collect(8, cap=1 << 18)
        assert ra == single and rb == single[3:8] and rc == single[::-1], rep


def test_nms_type_none_still_prunes(oracle_port):
    # ACF.cpp:334-351: with setDoNonMaximaSuppression(true) and pNms.type 'none', bbNms returns the boxes unchanged but
    # ObjectDetector::prune still cuts to m_maxDetectionCount / the score ratio
    opts = dict(small_face_opts(), nms_type="none")
    clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    det = _detector(opts, clf, True)
    img = synth.shapes_frame(3, 240, 320)
    rects, scores = det(img)
    assert len(rects) > 10
    det.setDoNonMaximaSuppression(True)
    det.setMaxDetectionCount(7)
    r2, s2 = det(img)
    kept = oracle_port.prune([(r[0], r[1], r[2], r[3], s) for r, s in zip(rects, scores)], 7, 0.0)
    assert [tuple(r) for r in r2] == [k[:4] for k in kept] and len(r2) == 7


def test_buffer_too_small_is_reported_by_collect():
    opts = small_face_opts()
    clf = synth.make_classifier(opts, 4, 2, seed=5, drift=1.0, gain=0.0, sigma=0.0)  # every window is a hit
    det = _detector(opts, clf, True, rows=256, cols=256, max_batch=1)
    img = synth.noise_frame(1, 96, 128)[None]
    det.submit(img.ctypes.data, 1, 96, 128, False)
    with pytest.raises(acf_b200.AcfError, match="too small"):
        det.collect(1, cap=8)


def test_nv12_frames_equal_host_conversion_then_rgb(oracle_port):
    # acfb_set_input_format(NV12): the device conversion must give exactly the pyramid / detections of converting on the host
    # with cv::cvtColor(COLOR_YUV2RGB_NV12)'s formula (synth.nv12_to_rgb) and handing the RGB24 frame over -- for random bytes
    # (every code path of the clamps) and for frames derived from the synthetic shapes, gray and LUV models
    for opts_fn, rows, cols in ((small_face_opts, 240, 320), (small_inria_opts, 202, 266)):
        opts = opts_fn()
        clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
        det = _detector(opts, clf, True, rows=256, cols=384, max_batch=2)
        rng = np.random.default_rng(11)
        for nv in (rng.integers(0, 256, (rows * 3 // 2, cols), dtype=np.uint8), synth.rgb_to_nv12(synth.shapes_frame(4, rows, cols))):
            rgb = synth.nv12_to_rgb(nv)
            det.setInputFormat("rgb")
            want_p, want_d = det.computePyramid(rgb), det(rgb)
            det.setInputFormat("nv12")
            got_p, got_d = det.computePyramid(nv), det(nv)
            for a, b in zip(got_p.data, want_p.data):
                assert np.array_equal(a, b)
            assert got_d == want_d
            Po = oracle_port.pyramid(opts, rgb)
            for a, b in zip(got_p.data, Po.data):
                assert np.array_equal(a, b)
        batch = np.stack([synth.rgb_to_nv12(synth.shapes_frame(s, rows, cols)) for s in (1, 2)])
        res = det(batch)
        assert res[1] == det(batch[1])
    with pytest.raises(acf_b200.AcfError, match="even"):
        det.computePyramid(np.zeros((96, 65), np.uint8))  # 64 rows x 65 cols: chroma pairs need even sizes


def test_deviation_from_the_reference_as_shipped_is_small_and_reported(oracle_ref_native):
    # SURVEY 8c protocol (3): parity is claimed against the exact-math build; against the SHIPPED build (SSE rcpps / rsqrtps,
    # channels off by up to ~5e-4) the hit sets may differ where a feature sits within that distance of a threshold.  Count them.
    import tools.native_deviation as nd
    opts = synth.face_opts(80)
    clf = synth.make_classifier(opts, 2048, 2, seed=1, n_reject=54)
    det = _detector(opts, clf, True, rows=1080, cols=1920, max_batch=1, cap=1 << 17)
    tot = dict(gpu=0, native=0, diff=0)
    for seed in (100, 102):
        r = nd.compare(det, oracle_ref_native, opts, clf, synth.shapes_frame(seed, 1080, 1920))
        print(seed, r)
        assert r["max_channel_delta"] < 0.1, "channel deviation from the shipped build beyond what its rcpps / rsqrtps explain (O off by up to 2e-2 rad near cos = +-1, SURVEY 8c)"
        assert r["max_score_delta"] < 0.5
        tot["gpu"] += r["gpu"]; tot["native"] += r["native"]; tot["diff"] += r["only_gpu"] + r["only_native"]
    assert tot["gpu"] > 0 and tot["native"] > 0
    assert tot["diff"] <= 0.25 * (tot["gpu"] + tot["native"]), tot


def _post_detector(opts, clf, device_post, rows=512, cols=640, max_batch=4, cap=1 << 16):
    os.environ["ACFB_DEVICE_POST"] = "1" if device_post else "0"
    try:
        det = _detector(opts, clf, True, rows=rows, cols=cols, max_batch=max_batch, cap=cap)
    finally:
        del os.environ["ACFB_DEVICE_POST"]
    return det


@pytest.mark.parametrize("nms_type,ovr_dnm,max_det,ratio", [
    ("maxg", "min", 7, 0.1), ("maxg", "union", 10, 0.0), ("max", "union", 10, 0.0), ("max", "min", 3, 0.5),
    ("maxg", "min", 64, 0.0), ("maxg", "union", 1, 0.0),
])
def test_device_side_order_rescale_nms_prune_equal_the_host_tail_and_the_oracle(oracle_port, nms_type, ovr_dnm, max_det, ratio):
    # k_post (post.cu): scale-major ordering, ACF.cpp:302-311 rescale, bbNms max / maxg (bbNms.cpp:111-192) and
    # ObjectDetector::prune (ObjectDetector.cpp:28-44) on the device == the engine's host tail == the oracle, box for box
    opts = dict(small_face_opts(), nms_type=nms_type, nms_ovrDnm=ovr_dnm, nms_overlap=0.4)
    clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    dev, host = _post_detector(opts, clf, True), _post_detector(opts, clf, False)
    fr = synth.frames("shapes", 4, 240, 320, seed0=3)
    raw = host(fr)
    assert min(len(r) for r, _ in raw) > 10
    for d in (dev, host):
        d.setDoNonMaximaSuppression(True)
        d.setMaxDetectionCount(max_det)
        d.setDetectionScorePruneRatio(ratio)
    rd, rh = dev(fr), host(fr)
    assert rd == rh
    for i in range(4):
        boxes = [(r[0], r[1], r[2], r[3], s) for r, s in zip(*raw[i])]
        kept = oracle_port.prune(oracle_port.nms(boxes, overlap=0.4, greedy=nms_type == "maxg", ovr_union=ovr_dnm == "union"), max_det, ratio)
        assert [tuple(r) for r in rd[i][0]] == [k[:4] for k in kept], i
        assert np.array_equal(np.array(rd[i][1], np.float32), np.array([k[4] for k in kept], np.float32)), i
    # batches in flight keep their order through the device tail as well
    a, b = np.ascontiguousarray(fr), np.ascontiguousarray(fr[::-1])
    dev.submit(a.ctypes.data, 4, 240, 320, False)
    dev.submit(b.ctypes.data, 4, 240, 320, False)
    ra, _ = dev.collect(4)
    rb, _ = dev.collect(4)
    assert ra == rd and rb == rd[::-1]


def test_device_tail_hands_hit_dense_frames_back_to_the_host(oracle_port):
    # more raw hits in a frame than k_post's shared-memory sort holds (4096): the batch falls back to the host tail
    # and the result is the same
    opts = dict(small_face_opts(), nms_type="maxg", nms_ovrDnm="min", nms_overlap=0.65)
    clf = synth.make_classifier(opts, 8, 2, seed=5, drift=1.0, gain=0.0, sigma=0.0)  # every window is a hit
    dev, host = _post_detector(opts, clf, True), _post_detector(opts, clf, False)
    fr = np.stack([synth.noise_frame(1, 240, 320), synth.shapes_frame(2, 240, 320)])
    for d in (dev, host):
        d.setDoNonMaximaSuppression(True)
    rd, rh = dev(fr), host(fr)
    assert rd == rh and all(len(r) >= 1 for r, _ in rd)
