"""torchrun helper of tests/test_gpu_dist.py: one process per GPU, the unique id travels over torch.distributed (plumbing),
the boxes over the engine's own NCCL communicator; rank 0 checks them against every rank's own acfb_collect result."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acf_b200  # noqa: E402
from acf_b200 import synth  # noqa: E402
from tests.test_gpu_dist import _as_lists, _detector  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for dense in (False, True):
    det = _detector(local, dense)
    fr = synth.frames("shapes", 4, 240, 320, seed0=40 + 10 * rank)
    want = det(fr)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(acf_b200.Detector.dist_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    det.dist_init_rank(bytes(uid.cpu().numpy().tobytes()), rank, world)
    for rep in range(3):
        det.submit(fr.ctypes.data, 4, 240, 320, False)
    got = []
    for rep in range(3):
        d, c, total = det.dist_collect_arrays(4)
        got.append(_as_lists(d.copy(), c.copy()))
    blob = [None] * world
    dist.all_gather_object(blob, pickle.dumps(want))
    if rank == 0:
        wants = [pickle.loads(b) for b in blob]
        for g in got:
            assert len(g) == 4 * world
            for r in range(world):
                for f in range(4):
                    assert [x[:4] for x in g[r * 4 + f]] == [tuple(x) for x in wants[r][f][0]], (dense, r, f)
                    assert [x[4] for x in g[r * 4 + f]] == wants[r][f][1]
                    assert all(x[5] == r * 4 + f for x in g[r * 4 + f])
        assert sum(len(x) for x in got[0]) > 0
    else:
        assert all(len(g) == 0 for g in got)
    det.close()
dist.barrier()
if rank == 0:
    print("dist ok")
dist.destroy_process_group()
