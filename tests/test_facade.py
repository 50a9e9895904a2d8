"""The header-only C++ facade (include/acf/ACF.h): compiles with plain g++ against the C ABI, reports
failure through good() like the reference (ACF.h:65-66), and -- on a GPU -- returns the same boxes as
the Python host mirror."""
import os
import subprocess

import numpy as np
import pytest

import acf_b200
from acf_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    out = tmp_path_factory.mktemp("facade") / "facade_demo"
    libdir = os.path.join(ROOT, "acf_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"), "-o", str(out),
                           "-L" + libdir, "-lacf_b200", "-Wl,-rpath," + libdir])
    return str(out)


def test_facade_compiles_and_reports_bad_model(demo):
    r = subprocess.run([demo, "/nonexistent/model.cpb", "x", "4", "4"], capture_output=True, text=True)
    assert r.returncode == 2 and "not good" in r.stderr


@pytest.mark.gpu
def test_facade_matches_python_mirror(demo, tmp_path):
    opts = synth.face_opts(32)
    clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    m = acf_b200.Model.create(opts, clf)
    m.save(tmp_path / "m.cpb")
    img = synth.shapes_frame(3, 240, 320)
    img.tofile(tmp_path / "f.rgb")
    det = acf_b200.Detector(m, max_rows=240, max_cols=320)
    det.setHitCapacity(1 << 17)
    for nms in (False, True):
        det.setDoNonMaximaSuppression(nms)
        det.setMaxDetectionCount(20)
        rects, scores = det(img, cap=1 << 18)
        args = [demo, str(tmp_path / "m.cpb"), str(tmp_path / "f.rgb"), "240", "320"] + (["nms"] if nms else [])
        env = dict(os.environ)
        r = subprocess.run(args, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        n, nscales = map(int, lines[0].split())
        assert n == len(rects) and nscales == len(det.plan(240, 320)[0])
        got = [tuple(map(float, l.split())) for l in lines[1:]]
        assert [tuple(int(v) for v in g[:4]) for g in got] == [tuple(r_) for r_ in rects]
        assert np.allclose([g[4] for g in got], scores, rtol=0, atol=1e-6)
