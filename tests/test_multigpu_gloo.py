"""CPU, world_size 2 on gloo: the N>1 host logic -- contiguous batch sharding and the gather of the
variable-length detection lists on rank 0 (the only exchange of the path)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from acf_b200 import dist as adist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_results(lo, hi):
    rng = np.random.default_rng(7)
    alln = rng.integers(0, 5, 64)
    out = []
    for f in range(lo, hi):
        k = int(alln[f])
        out.append(([(f, j, 10 + j, 20 + j) for j in range(k)], [float(f) + 0.25 * j for j in range(k)]))
    return out


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = adist.shard_range(n_frames, world, rank)
    res = adist.gather_detections(_fake_results(lo, hi), dist, "cpu", frame0=lo)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    for n in (1, 7, 64, 2048):
        for world in (1, 2, 4, 8):
            r = [adist.shard_range(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1


def test_gather_of_detection_lists_world2():
    world, n_frames = 2, 13  # ragged shards: 6 + 7 frames, some frames empty
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_results(0, n_frames)
    assert len(got) == n_frames
    for g, w in zip(got, want):
        assert g[0] == w[0] and np.allclose(g[1], w[1])


def _fake_arrays(lo, hi):
    from acf_b200.detector import DET_DTYPE
    res = _fake_results(lo, hi)
    counts = np.array([len(r[0]) for r in res], np.int32)
    dets = np.zeros(int(counts.sum()), DET_DTYPE)
    k = 0
    for f, (rects, scores) in enumerate(res):
        for rc, s in zip(rects, scores):
            dets[k] = (rc[0], rc[1], rc[2], rc[3], s, f)  # frame index local to the rank, as acfb_collect reports it
            k += 1
    return dets, counts


def _worker_arrays(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = n_frames // world
    dets, counts = _fake_arrays(rank * per, (rank + 1) * per)
    res = adist.gather_detection_arrays(dets, counts, dist, "cpu", frame0=rank * per)
    if rank == 0:
        q.put((res[0].tolist(), res[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_of_detection_arrays_world2():
    # the form bench.py uses: structured arrays from Detector.collect_arrays, equal frame counts per rank (batch sharding)
    world, n_frames = 2, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_arrays, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in ps:
        p.start()
    dets, counts = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_results(0, n_frames)
    assert counts == [len(w[0]) for w in want]
    k = 0
    for f, (rects, scores) in enumerate(want):
        for rc, s in zip(rects, scores):
            x, y, w, h, sc, fr = dets[k]
            assert (x, y, w, h) == rc and abs(sc - s) < 1e-6 and fr == f
            k += 1
    assert k == len(dets)


def test_shared_memory_exchange_of_the_engine_on_the_host():
    """the engine's single-node exchange (csrc/dist.cpp ShmExchange, what acfb_dist_collect uses with ACFB_DIST_EXCHANGE=shm or
    without NCCL) needs no device: ranks as threads, 40 batches through a ring of 8 (wrap-around + flow control), every record
    byte-checked on rank 0"""
    from acf_b200 import _capi
    for world in (1, 2, 8):
        _capi.check(_capi.lib().acfb_selftest_exchange(world, 40, 4096))
    assert _capi.lib().acfb_selftest_exchange(0, 1, 4096) != 0
