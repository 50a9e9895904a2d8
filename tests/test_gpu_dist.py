"""Multi-GPU gather inside the engine (acfb_dist_*): rank 0 must receive exactly the boxes every rank's own
acfb_collect returns, in global frame order -- with both exchanges (shared-memory ring, the single-node default; ncclAllGather of
the device buffer k_post wrote, ACFB_DIST_EXCHANGE=nccl), for batches finished by k_post and for batches the host tail had to
finish (a hit-dense frame k_post hands back).  World size 1 runs on any GPU box; the
two-device cases (one process / two host threads over ncclCommInitAll, and two processes over a broadcast unique id) need a
box with two GPUs (gpurun --gpus 2) and are skipped elsewhere."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import acf_b200
from acf_b200 import synth
from tests.golden.make_golden import small_face_opts

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _detector(device, dense=False, max_batch=4):
    opts = dict(small_face_opts(), nms_type="maxg", nms_ovrDnm="min", nms_overlap=0.5)
    if dense:  # every window is a hit: more raw hits per frame than k_post sorts -> the batch falls back to the host tail
        clf = synth.make_classifier(opts, 8, 2, seed=5, drift=1.0, gain=0.0, sigma=0.0)
    else:
        clf = synth.make_classifier(opts, 64, 2, seed=5, drift=-0.05, gain=0.3)
    det = acf_b200.Detector(acf_b200.Model.create(opts, clf), device=device, max_rows=256, max_cols=320, max_batch=max_batch)
    det.setHitCapacity(1 << 16)
    det.setDoNonMaximaSuppression(True)
    det.setMaxDetectionCount(7)
    return det


def _as_lists(dets, counts):
    out, k = [], 0
    for c in counts:
        out.append([(int(d["x"]), int(d["y"]), int(d["w"]), int(d["h"]), float(d["score"]), int(d["frame"])) for d in dets[k:k + c]])
        k += c
    return out


@pytest.fixture(params=["shm", "nccl"])
def exchange(request):
    old = os.environ.get("ACFB_DIST_EXCHANGE")
    os.environ["ACFB_DIST_EXCHANGE"] = request.param
    yield request.param
    if old is None:
        del os.environ["ACFB_DIST_EXCHANGE"]
    else:
        os.environ["ACFB_DIST_EXCHANGE"] = old


@pytest.mark.parametrize("dense", [False, True])
def test_world_of_one_equals_collect(dense, exchange):
    det = _detector(0, dense)
    fr = synth.frames("shapes", 4, 240, 320, seed0=11)
    want = det(fr)
    det.dist_init_rank(acf_b200.Detector.dist_unique_id(), 0, 1)
    assert det.dist_info()[:2] == (0, 1) and (det.dist_info()[2] > 20000) == (exchange == "nccl")
    for rep in range(2):  # two batches in flight
        det.submit(fr.ctypes.data, 4, 240, 320, False)
    for rep in range(2):
        dets, counts, total = det.dist_collect_arrays(4)
        got = _as_lists(dets, counts)
        assert total == sum(len(r) for r, _ in want) > 0
        for f in range(4):
            assert [g[:4] for g in got[f]] == [tuple(r) for r in want[f][0]]
            assert [g[4] for g in got[f]] == want[f][1] and all(g[5] == f for g in got[f])


@pytest.mark.parametrize("dense", [False, True])
def test_one_process_two_devices_two_threads(dense, exchange):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    dets = [_detector(0, dense), _detector(1, dense)]
    frames = [synth.frames("shapes", 4, 240, 320, seed0=20), synth.frames("shapes", 4, 240, 320, seed0=30)]
    want = [d(f) for d, f in zip(dets, frames)]
    acf_b200.Detector.dist_init_all(dets)
    out = [None, None]

    def run(r):
        res = []
        for rep in range(3):
            dets[r].submit(frames[r].ctypes.data, 4, 240, 320, False)
        for rep in range(3):
            res.append(dets[r].dist_collect_arrays(4))
            res[-1] = (res[-1][0].copy(), res[-1][1].copy(), res[-1][2])
        out[r] = res

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert out[0] is not None and out[1] is not None, "a rank did not finish"
    for d, c, total in out[1]:
        assert total == 0 and len(d) == 0
    for d, c, total in out[0]:
        got = _as_lists(d, c)
        assert len(got) == 8
        for r in range(2):
            for f in range(4):
                assert [g[:4] for g in got[r * 4 + f]] == [tuple(x) for x in want[r][f][0]], (r, f)
                assert [g[4] for g in got[r * 4 + f]] == want[r][f][1]
                assert all(g[5] == r * 4 + f for g in got[r * 4 + f])


def test_two_processes_over_a_broadcast_unique_id(exchange):
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_two_ranks.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dist ok" in r.stdout
