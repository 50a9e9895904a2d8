"""Multi-GPU plumbing: frames shard by batch across ranks (one process per GPU); the only exchange is the
gather of the variable-length per-frame detection lists on rank 0.  torch.distributed is
plumbing only (NCCL on GPUs, gloo in the CPU tests); no kernel of the data path is involved."""
import numpy as np

DET_WORDS = 6  # x, y, w, h, score, frame  (24 bytes, acfb_det)


def shard_range(n_frames, world, rank):
    """contiguous shard [lo, hi) of rank `rank`: GPU g gets frames [g*N/G, (g+1)*N/G)"""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def pack_detections(results, frame0=0, cap=64):
    """results: per-frame (rects, scores) -> (counts[int32 n], payload[float32 n, cap, 6]); boxes beyond cap per
    frame are dropped from the payload but still counted."""
    n = len(results)
    counts = np.zeros(n, np.int32)
    pay = np.zeros((n, cap, DET_WORDS), np.float32)
    for f, (rects, scores) in enumerate(results):
        counts[f] = len(rects)
        for j, (rc, s) in enumerate(list(zip(rects, scores))[:cap]):
            pay[f, j] = (rc[0], rc[1], rc[2], rc[3], s, frame0 + f)
    return counts, pay


def gather_detections(results, dist, device="cpu", frame0=0, cap=64):
    """every rank calls this with its own per-frame results; rank 0 gets [(rects, scores)] for all frames of
    all ranks in global frame order, the others get None.  Two collectives: counts, then fixed-capacity payload."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts, pay = pack_detections(results, frame0, cap)
    tc = torch.from_numpy(counts).to(device)
    tp = torch.from_numpy(pay).to(device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(results)], dtype=torch.int64, device=device))
    nmax = int(max(int(s.item()) for s in sizes))
    pc = torch.zeros(nmax, dtype=torch.int32, device=device); pc[:len(results)] = tc
    pp = torch.zeros((nmax, cap, DET_WORDS), dtype=torch.float32, device=device); pp[:len(results)] = tp
    all_c = [torch.empty_like(pc) for _ in range(world)] if rank == 0 else None
    all_p = [torch.empty_like(pp) for _ in range(world)] if rank == 0 else None
    dist.gather(pc, all_c, dst=0)
    dist.gather(pp, all_p, dst=0)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        n = int(sizes[r].item())
        c = all_c[r].cpu().numpy(); p = all_p[r].cpu().numpy()
        for f in range(n):
            k = min(int(c[f]), cap)
            rects = [tuple(int(v) for v in p[f, j, :4]) for j in range(k)]
            scores = [float(p[f, j, 4]) for j in range(k)]
            out.append((rects, scores))
    return out


def gather_detection_arrays(dets, counts, dist, device="cpu", frame0=0):
    """array form of gather_detections for Detector.collect_arrays: dets = structured array (x, y, w, h, score, frame) of this
    rank's batch, counts = int32 per frame.  Two collectives (per-frame counts + total, then the records padded to the
    largest rank), one host synchronisation, no per-detection Python objects.  Rank 0 gets (dets of all ranks in global frame
    order with `frame` made global, counts of all frames); the other ranks get None."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(counts)
    head = np.empty(n + 1, np.int32)
    head[:n] = counts; head[n] = len(dets)
    th = torch.from_numpy(head).to(device)
    heads = torch.empty(world * (n + 1), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(heads, th)
    heads = heads.cpu().numpy().reshape(world, n + 1)  # the one synchronisation: every rank needs the padded record count
    if not np.all(heads[:, :n].sum(axis=1) == heads[:, n]):
        raise RuntimeError("gather_detection_arrays: ranks disagree on the frames per rank")
    tmax = int(heads[:, n].max())
    pay = np.zeros((max(1, tmax), DET_WORDS), np.int32)
    if len(dets):
        rec = np.ascontiguousarray(dets).view(np.int32).reshape(len(dets), DET_WORDS).copy()
        rec[:, 5] += frame0
        pay[:len(dets)] = rec
    tp = torch.from_numpy(pay).to(device)
    allp = [torch.empty_like(tp) for _ in range(world)] if rank == 0 else None
    dist.gather(tp, allp, dst=0)
    if rank != 0:
        return None
    from .detector import DET_DTYPE
    parts = [allp[r].cpu().numpy()[: int(heads[r, n])] for r in range(world)]
    return np.concatenate(parts).view(DET_DTYPE).reshape(-1), heads[:, :n].reshape(-1).copy()
