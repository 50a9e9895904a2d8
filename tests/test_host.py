"""CPU: host-side logic of the product -- the C ABI loads and exports every declared symbol, the
.cpb reader/writer round-trips, model validation and acfModify behave like the reference.  No
kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import acf_b200
from acf_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "acf_b200.h")).read()
    declared = set(re.findall(r"ACFB_API[^;(]*?\b(acfb_\w+)\s*\(", header))
    assert declared, "no declarations found"
    L = _capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/acf_b200.h but not exported"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    assert b"sm_100a" in L.acfb_version()


def _model(opts=None, n_trees=32, depth=2):
    opts = opts or synth.face_opts(64)
    return acf_b200.Model.create(opts, synth.make_classifier(opts, n_trees, depth, seed=3)), opts


def test_cpb_round_trip_keeps_tables_bit_equal(tmp_path):
    # mirrors the reference's cpb test (src/test/test-acf-api.cpp:465-478,604-642): fids/child/depth equal
    m, opts = _model(synth.inria_opts())
    blob = m.to_bytes()
    assert blob[0] == 1  # cereal PortableBinary endianness flag
    m2 = acf_b200.Model.load(blob)
    c1, c2 = m.classifier, m2.classifier
    for k in ("fids", "child", "depth", "thrs", "hs"):
        assert np.array_equal(c1[k], c2[k]), k
    assert c1["treeDepth"] == c2["treeDepth"] == 2
    o2 = m2.options
    for k in ("shrink", "colorSpace", "pad", "modelDs", "modelDsPad", "lambdas", "stride", "cascThr", "nPerOct", "nApprox"):
        assert o2[k] == (tuple(opts[k]) if isinstance(opts[k], tuple) else opts[k]), k
    assert m2.to_bytes() == blob  # writer and reader are symmetric
    p = tmp_path / "synthetic_seed3.cpb"
    m.save(p)
    assert acf_b200.Model.load(str(p)).to_bytes() == blob


def test_cpb_layout_follows_cereal_rules():
    m, _ = _model()
    b = m.to_bytes()
    # u8 endian | u32 Detector version (=1) | u32 Classifier version | u32 cv::Mat version | rows, cols, type
    assert b[0] == 1
    v_det, v_clf, v_mat, rows, cols, typ = np.frombuffer(b[1:25], np.uint32)
    assert (v_det, v_clf, v_mat) == (1, 0, 0) and (rows, cols, typ) == (32, 7, 4)  # CV_32S fids
    assert b[25] == 1  # continuous flag


def test_cpb_rejects_garbage_and_truncation():
    m, _ = _model()
    b = m.to_bytes()
    with pytest.raises(acf_b200.AcfError):
        acf_b200.Model.load(b[: len(b) // 2])
    with pytest.raises(acf_b200.AcfError):
        acf_b200.Model.load(b + b"\0")
    with pytest.raises(acf_b200.AcfError):
        acf_b200.Model.load(b"\x07" + b[1:])
    with pytest.raises(acf_b200.AcfError):
        acf_b200.Model.load("/nonexistent/model.mat")  # only .cpb is accepted (ACFIO.cpp:204-208)


def test_model_validation_rejects_bad_feature_ids():
    opts = synth.face_opts(64)
    clf = synth.make_classifier(opts, 8, 2, seed=0)
    clf["fids"][0, 0] = 10 ** 6
    with pytest.raises(acf_b200.AcfError):
        acf_b200.Model.create(opts, clf)


def test_acf_modify_is_cumulative_and_rerounds_stride():
    # acfModify.cpp:139,143
    m, _ = _model()
    hs0 = m.classifier["hs"].copy()
    m.acfModify(cascCal=0.25)
    m.acfModify(cascCal=0.25, stride=6)
    assert np.allclose(m.classifier["hs"], hs0 + 0.5, atol=1e-6)
    assert m.options["stride"] == 8  # round(6/4)*4
    m.acfModify(cascThr=-2.0)
    assert m.options["cascThr"] == -2.0


def test_engine_creation_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, _ = _model()
    with pytest.raises(acf_b200.AcfError, match="no CPU fallback|CUDA"):
        acf_b200.Detector(m, max_rows=64, max_cols=64)


def test_synthetic_frames_are_deterministic():
    a, b = synth.shapes_frame(5, 120, 160), synth.shapes_frame(5, 120, 160)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (120, 160, 3)
    assert not np.array_equal(a, synth.shapes_frame(6, 120, 160))
    n = synth.noise_frame(1, 64, 96)
    assert n.std() > 5


def test_get_scales_matches_oracle_over_many_sizes(oracle_port):
    # the scale schedule drives every buffer size and box coordinate: the product's host restatement
    # (acf_b200/csrc/plan.cpp) against the oracle's, bit for bit, over frame sizes and model shapes
    rng = np.random.default_rng(4)
    sizes = [(1080, 1920), (2160, 3840), (480, 640), (640, 480), (720, 1280), (128, 160)]
    sizes += [(int(r) * 4, int(c) * 4) for r, c in rng.integers(20, 400, (12, 2))]
    for opts in (synth.face_opts(80), synth.face_opts(64), synth.inria_opts(), dict(synth.face_opts(32), nPerOct=4, nOctUp=1)):
        for rows, cols in sizes:
            s, hw = acf_b200.get_scales(opts, rows, cols)
            so, hwo = oracle_port.get_scales(opts["nPerOct"], opts["nOctUp"], opts["minDs"], opts["shrink"], (rows, cols))
            assert np.array_equal(s, so) and np.array_equal(hw, hwo), (rows, cols)


def test_operator_entry_points_reject_a_missing_engine():
    """acfb_op_* run on a GPU engine; without one they must fail with an error code, not compute anything on the host."""
    import ctypes as C
    L = acf_b200.lib()
    buf = np.zeros(64, np.float32)
    n = C.c_int(0)
    assert L.acfb_op_rgb_convert(None, buf.ctypes.data, 4, 4, 0, buf.ctypes.data, C.byref(n)) != 0
    assert L.acfb_op_conv_tri(None, buf.ctypes.data, 8, 8, 1, 1.0, buf.ctypes.data) != 0
    assert L.acfb_op_gradient_mag(None, buf.ctypes.data, 8, 8, 1, 0, 0, 0.005, 0, buf.ctypes.data, None) != 0
    assert L.acfb_op_gradient_hist(None, buf.ctypes.data, buf.ctypes.data, 8, 8, 4, 6, 0, 0, 0.2, 0, buf.ctypes.data) != 0
    assert L.acfb_op_im_resample(None, buf.ctypes.data, 8, 8, 1, 4, 4, 1.0, buf.ctypes.data) != 0
    assert b"null" in L.acfb_last_error()


def test_cpb_loader_survives_mutated_archives(tmp_path):
    """A corrupt length field must be refused before anything is allocated for it (a mutated cv::Mat type once asked for
    a terabyte), and no mutation may crash or stall the loader: 600 random byte / word mutations and truncations."""
    import ctypes as C
    import time
    opts = synth.face_opts(32)
    m = acf_b200.Model.create(opts, synth.make_classifier(opts, 8, 2, seed=1))
    path = tmp_path / "m.cpb"
    m.save(path)
    good = bytearray(open(path, "rb").read())
    L = acf_b200.lib()
    rng = np.random.default_rng(7)
    t0 = time.time()
    loaded = 0
    for _ in range(600):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, len(b)))
            width = (1, 4, 8)[int(rng.integers(0, 3))]
            b[pos:pos + width] = int(rng.integers(0, 2 ** 63)).to_bytes(8, "little")[:width]
        if rng.random() < 0.2:
            b = b[: int(rng.integers(1, len(b)))]
        arr = (C.c_ubyte * len(b)).from_buffer(b)
        h = C.c_void_p()
        if L.acfb_model_load(arr, len(b), C.byref(h)) == 0:
            loaded += 1
            L.acfb_model_destroy(h)
    assert time.time() - t0 < 60
    assert 0 < loaded < 600  # mutations of thresholds / leaf values still load; structural ones are refused


def test_options_that_drive_host_loops_are_validated():
    opts = synth.face_opts(32)
    clf = synth.make_classifier(opts, 8, 2, seed=1)
    for bad in (dict(nPerOct=0), dict(nPerOct=1 << 20), dict(nApprox=-1), dict(nOctUp=-1), dict(minDs=(0, 32)), dict(pad=(-4, 0))):
        with pytest.raises(RuntimeError):
            acf_b200.Model.create(dict(opts, **bad), clf)
    with pytest.raises(RuntimeError):
        acf_b200.get_scales(dict(opts, nPerOct=0), 240, 320)


def test_variable_depth_child_links_are_validated():
    # a corrupt treeDepth == 0 archive must be refused at load time -- a link that points backwards would make the
    # cascade walk forever, a link past the last node reads outside the tree's record
    opts = synth.face_opts(64)
    good = synth.make_variable_classifier(opts, 8, 3, seed=2)
    acf_b200.Model.create(opts, good)
    nn = good["child"].shape[1]
    for bad_link in (1, nn, nn + 5):
        clf = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in good.items()}
        clf["child"][0, 0] = bad_link  # 1: the root's own slot (loop); nn: right child outside; nn + 5: both outside
        with pytest.raises(acf_b200.AcfError, match="child link"):
            acf_b200.Model.create(opts, clf)
    clf = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in good.items()}
    clf["child"][3, 2] = 2  # points at an earlier node
    with pytest.raises(acf_b200.AcfError, match="child link"):
        acf_b200.Model.create(opts, clf)
    # the same through the archive: flip the link inside the serialised bytes
    blob = bytearray(acf_b200.Model.create(opts, good).to_bytes())
    child = np.ascontiguousarray(good["child"], np.uint32).tobytes()
    at = bytes(blob).find(child)
    assert at > 0
    blob[at:at + 4] = np.uint32(1).tobytes()
    with pytest.raises(acf_b200.AcfError, match="child link"):
        acf_b200.Model.load(bytes(blob))


@pytest.mark.parametrize("name,opts_fn,depth", [("FACE80", lambda: synth.face_opts(80), 2), ("FACE80c", lambda: synth.face_opts(80, True), 2),
                                                ("INRIA", synth.inria_opts, 2), ("FACE64-depth4", lambda: synth.face_opts(64), 4)])
def test_cpb_bytes_match_independent_packer(name, opts_fn, depth):
    # reader and writer share one walker, so their round trip proves symmetry only.  tests/cpb_pack.py is a second
    # writer, built from the reference's serialisers (ACFIOArchive.h:75-216, io/cvmat_cereal.h:18-46, ACFField.h:123-130) and
    # cereal's rules, that never looks at cpb.cpp: field order, widths, Field<T> names / has / isLeaf and the once-only
    # class-version words must agree byte for byte, and the engine's reader must load what it writes.
    from tests import cpb_pack
    opts = opts_fn()
    clf = synth.make_classifier(opts, 24, depth, seed=4)
    mine = acf_b200.Model.create(opts, clf).to_bytes()
    ref = cpb_pack.pack(opts, clf, weights=clf["weights"], depth=clf["depth"])
    if mine != ref:
        at = next(i for i, (a, b) in enumerate(zip(mine, ref)) if a != b) if len(mine) == len(ref) or True else -1
        raise AssertionError(f"{name}: archives differ at byte {at} of {len(mine)} / {len(ref)}: ours {mine[at:at + 16].hex()} packer {ref[at:at + 16].hex()}")
    m2 = acf_b200.Model.load(ref)
    assert m2.to_bytes() == ref
    assert np.array_equal(m2.classifier["fids"], clf["fids"]) and np.array_equal(m2.classifier["hs"], clf["hs"])


def test_nv12_definition_matches_opencv():
    # the NV12 -> RGB8 conversion the engine runs on the device is DEFINED as cv::cvtColor(COLOR_YUV2RGB_NV12)'s integer formula
    # (acf_b200/synth.py:nv12_to_rgb restates it; the GPU test compares the device against that function)
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for rows, cols in ((8, 12), (270, 480)):
        nv = rng.integers(0, 256, (rows * 3 // 2, cols), dtype=np.uint8)
        assert np.array_equal(synth.nv12_to_rgb(nv), cv2.cvtColor(nv, cv2.COLOR_YUV2RGB_NV12))
    rgb = synth.noise_frame(3, 64, 96)
    back = synth.nv12_to_rgb(synth.rgb_to_nv12(rgb)).astype(int)
    assert np.abs(back - rgb.astype(int)).mean() < 6  # chroma subsampling: a sanity bound on the round trip, nothing more


def test_acf_modify_merges_the_whole_modify_field_set():
    # Detector::Modify (ACF.h:392-408) merged field by field (acfModify.cpp:99-123): only the given groups change, the scale
    # schedule follows nPerOct / nOctUp / minDs / pad, a model the engine could not run is refused and left untouched
    m, opts = _model()
    before = m.options
    hs0 = m.classifier["hs"].copy()
    m.acfModify(cascCal=0.125, nPerOct=4, nOctUp=1, nApprox=3, lambdas=[0.0, 0.11, 0.12], pad=(8, 8), minDs=(48, 48),
                pNms=dict(type="max", overlap=0.5, ovrDnm="union"), cascThr=-2.0, stride=6)
    o = m.options
    assert (o["nPerOct"], o["nOctUp"], o["nApprox"]) == (4, 1, 3)
    assert list(o["lambdas"]) == [0.0, 0.11, 0.12] and tuple(o["pad"]) == (8, 8) and tuple(o["minDs"]) == (48, 48)
    assert (o["nms_type"], o["nms_overlap"], o["nms_ovrDnm"]) == ("max", 0.5, "union")
    assert o["cascThr"] == -2.0 and o["stride"] == 8 and o["shrink"] == before["shrink"]  # 6 -> round(6 / 4) * 4
    assert np.array_equal(m.classifier["hs"], (hs0.astype(np.float64) + 0.125).astype(np.float32))
    s_after, _ = acf_b200.get_scales(o, 480, 640)
    s_before, _ = acf_b200.get_scales(before, 480, 640)
    assert len(s_after) != len(s_before) and s_after[0] == 2.0  # one octave up, four scales per octave
    # only one group given: everything else stays
    m.acfModify(nApprox=0)
    o2 = m.options
    assert o2["nApprox"] == 0 and o2["nPerOct"] == 4 and tuple(o2["pad"]) == (8, 8) and o2["cascThr"] == -2.0
    # the round trip through the archive keeps the merged options
    o3 = acf_b200.Model.load(m.to_bytes()).options
    assert (o3["nPerOct"], o3["nOctUp"], o3["nApprox"], o3["nms_type"]) == (4, 1, 0, "max")
    with pytest.raises(acf_b200.AcfError):
        m.acfModify(nPerOct=0)
    assert m.options["nPerOct"] == 4
