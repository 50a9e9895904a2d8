#!/usr/bin/env python
"""bench.py -- 1080p frames/s and Mwindows/s of the chnsPyramid + acfDetect path on B200.

A "step" is one pass of the hot path over one batch of synthetic frames (BASELINE.json configs[1]:
1080p, batch 256 per GPU, FACE80-shaped 7-channel model, full 31-scale pyramid + cascade).
  value : frames/s with the u8 frames already resident in HBM (acfb_submit on_device=1 + acfb_collect)
  e2e   : the same through the public C ABI with HOST (pinned) frames -- H2D of every frame and D2H of
          the hit lists inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"
Multi-GPU (torchrun, one rank per GPU): frames shard by batch (weak scaling, 256 per GPU), no data-path
collective; the only exchange is the NCCL gather of the per-frame detection lists on rank 0.
--impl reference times the reference's own CPU implementation (oracle/_ref, native SSE arithmetic)
on the host cores for the same metric.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REAL_GROUP = ["k_resample", "k_resample_x", "k_resample_y", "k_smooth", "k_gradmag", "k_trix", "k_triyhist", "k_hist"]
WINDOWS_1080P_FACE80 = 662799


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--rows", type=int, default=1080)
    ap.add_argument("--cols", type=int, default=1920)
    ap.add_argument("--model", default="face80", choices=["face80", "face80c", "inria"])
    ap.add_argument("--distinct", type=int, default=16, help="distinct synthetic frames tiled to the batch")
    ap.add_argument("--trees", type=int, default=2048)
    ap.add_argument("--operating-point", default="hits", choices=["hits", "fast", "deep"],
                    help="synthetic cascade: hits (headline: ~12 trees/window, the survivors of 52 rejector trees walk all 2048 trees and "
                         "become ~200 raw hits per frame), fast (the same rejectors, no hits) or deep (~75 trees/window, no hits)")
    ap.add_argument("--input-format", default="rgb", choices=["rgb", "gray"],
                    help="frames handed to the detector: RGB24 (default, the headline) or GRAY8 (one third of the PCIe bytes; "
                         "gray / orig models only, chnsPyramid.cpp:234-244)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = 2 per core)")
    return ap.parse_args()


def model_opts(name):
    from acf_b200 import synth
    return {"face80": lambda: synth.face_opts(80), "face80c": lambda: synth.face_opts(80, True), "inria": synth.inria_opts}[name]()


def algorithmic_bytes(det, rows, cols, hits_per_frame):
    """B_pyr = in_u8 + chan_f32 ; B_det = chan_f32 + 24*hits ; chan_f32 = dense padded channel bytes."""
    info, _ = det.plan(rows, cols)
    chan = sum(s.nchn * s.w * s.h for s in info) * 4
    in_u8 = rows * cols * 3
    return dict(in_u8=in_u8, chan_f32=chan, B_pyr=in_u8 + chan, B_det=chan + 24 * hits_per_frame,
                B_frame=in_u8 + 2 * chan + 24 * hits_per_frame)


class ClockSampler:
    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # keep the samples taken under load (upper half) for the median
        sm_sorted = sorted(sm)
        med = float(np.median(sm_sorted[len(sm_sorted) // 2:])) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_run(opts, clf, frames, threads):
    """frame-parallel CPU path (mirrors src/app/acf/acf.cpp:443-455: one detector state per thread).
    Returns (seconds, frames, kind, hits)."""
    from oracle import oracle as O
    kind = "ref_native" if O.available("ref_native") else "port"
    orc = O.Oracle(kind)
    oo = O.opts_from_dict(opts)
    nf = len(frames)
    hits = [0] * nf

    def work(tid):
        for i in range(tid, nf, threads):
            P = orc.pyramid(oo, frames[i])
            _, _, _, total = P.detect(clf, cap=1 << 16)
            hits[i] = total
            P.close()
    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, nf, ("reference" if kind == "ref_native" else "port"), sum(hits)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from acf_b200 import synth
    opts = model_opts(a.model)
    clf = {"hits": lambda: synth.make_classifier(opts, a.trees, 2, seed=1, n_reject=52 if a.model != "inria" else None),
           "fast": lambda: synth.make_classifier(opts, a.trees, 2, seed=1, n_reject=a.trees if a.model == "inria" else None),
           "deep": lambda: synth.make_classifier(opts, a.trees, 2, seed=1, drift=-0.113, gain=0.23)}[a.operating_point]()
    windows_per_frame = None
    config = {"workload": f"{a.rows}x{a.cols} synthetic 'shapes' frames, batch {a.batch}/GPU, {a.model} "
                          f"({synth.n_channels(opts)} channels, {a.trees} depth-2 trees), full pyramid + cascade",
              "frames_per_step_per_gpu": a.batch, "distinct_frames": a.distinct, "model": a.model, "operating_point": a.operating_point,
              "l2_policy": f"inputs larger than L2 ({a.batch * a.rows * a.cols * (3 if a.input_format == 'rgb' else 1) / 1e9:.2f} GB of u8 frames per step per GPU)"}

    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        nf = a.cpu_frames or max(16, 4 * cores)
        frames = synth.frames("shapes", min(nf, a.distinct), a.rows, a.cols, seed0=100)
        frames = [frames[i % len(frames)] for i in range(nf)]
        for _ in range(max(0, min(a.warmup, 1))):
            cpu_run(opts, clf, frames[:cores], cores)
        t = 0.0; n = 0
        for _ in range(a.steps):
            dt, k, kind, _ = cpu_run(opts, clf, frames, cores)
            t += dt; n += k
        fps = n / t
        line = {"impl": "reference", "metric": "frames_per_sec_1080p", "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * t / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                 "sample": f"{nf} frames per step, frame-parallel over {cores} threads, native SSE arithmetic"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "mwindows_per_sec": fps * WINDOWS_1080P_FACE80 / 1e6 if (a.rows, a.cols) == (1080, 1920) else None}
        print(json.dumps(line))
        return

    import torch
    import acf_b200
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = acf_b200.Model.create(opts, clf)
    det = acf_b200.Detector(model, device=local, max_rows=a.rows, max_cols=a.cols, max_batch=a.batch)
    det.setHitCapacity(8192)
    info, _ = det.plan(a.rows, a.cols)
    # synthetic frames: `distinct` seeded frames tiled to the batch; different seeds per rank
    base = synth.frames("shapes", a.distinct, a.rows, a.cols, seed0=100 + 1000 * rank)
    bpp = 3 if a.input_format == "rgb" else 1
    if bpp == 1:
        det.setInputFormat("gray")
        config["input_format"] = "GRAY8 (green plane of the synthetic frames)"
    host = torch.empty((a.batch, a.rows, a.cols, bpp), dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for i in range(a.batch):
        hv[i] = base[i % a.distinct] if bpp == 3 else base[i % a.distinct][:, :, 1:2]
    dev = host.cuda(non_blocking=False)
    stream = torch.cuda.ExternalStream(det.stream(), device=local)
    cap = 1 << 18

    def step(on_device):
        det.submit(dev.data_ptr() if on_device else host.data_ptr(), a.batch, a.rows, a.cols, on_device)
        res, total = det.collect(a.batch, cap=cap)
        return res, total

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(res):
        """NCCL gather of the variable-length detection lists on rank 0 (counts, then fixed-capacity payload)"""
        if dist is None:
            return res
        from acf_b200 import dist as adist
        return adist.gather_detections(res, dist, "cuda", frame0=rank * a.batch)

    def timed(on_device, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = det.launch_count()
        t0 = time.perf_counter()
        e0.record(stream)
        tot_hits = 0
        stage_acc = {}
        # public asynchronous API with up to three batches in flight: while the host orders / rescales the hits of step k, the
        # kernels of step k+1 already run; with host frames the H2D copy of step k+1 (copy stream) overlaps the kernels
        # of step k -- every step still copies its own frames from pinned host memory inside the timed region
        ptr = dev.data_ptr() if on_device else host.data_ptr()
        stage_on = getattr(det, "_timing", False)
        if stage_on:
            for _ in range(steps):
                res, total = step(on_device)
                tot_hits += total
                for nme, ms in det.stage_times():
                    stage_acc[nme] = stage_acc.get(nme, 0.0) + ms
                gather(res)
        else:
            depth = 2  # batches submitted ahead of the one being collected (the engine keeps up to three in flight)
            for k in range(min(depth, steps)):
                det.submit(ptr, a.batch, a.rows, a.cols, on_device)
            for k in range(steps):
                if k + depth < steps:
                    det.submit(ptr, a.batch, a.rows, a.cols, on_device)
                res, total = det.collect(a.batch, cap=cap)
                tot_hits += total
                gather(res)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        ms = max(ms, 0.0)
        # the stream is idle while the host orders / rescales hits, so the device-event span covers the whole step
        t = torch.tensor([ms, wall * 1000.0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), tot_hits, det.launch_count() - l0, {k: v / steps for k, v in stage_acc.items()}

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # samples clocks / throttle reasons through warm-up and both timed regions
    for _ in range(max(3, a.warmup)):
        res, _ = step(True)
        gather(res)  # also warms up the NCCL communicator (lazy initialisation would land in the timed region)
    ms_dev, wall_dev, hits_dev, launches, _ = timed(True, a.steps)
    # per-kernel device times for the roofline: a second, instrumented pass over the same K steps with CUDA events
    # between the kernels on the engine's stream (instrumentation serialises the two compute streams the engine
    # otherwise overlaps, so these durations are per kernel, not per step)
    det.enable_stage_timing(True); det._timing = True
    step(True)
    _, _, _, _, stages = timed(True, a.steps)
    det.enable_stage_timing(False); det._timing = False
    timed(False, 3)  # untimed: three host batches in flight, so every slot's staging buffer exists before the timed region
    ms_e2e, wall_e2e, hits_e2e, _, stages_e2e = timed(False, a.steps)
    _, trees, windows = det.last_hits()
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total_frames = a.batch * world * a.steps
    t_dev = max(ms_dev, wall_dev) / 1000.0
    t_e2e = max(ms_e2e, wall_e2e) / 1000.0
    fps = total_frames / t_dev
    fps_e2e = total_frames / t_e2e
    windows_per_frame = windows / a.batch
    hits_per_frame = hits_dev / (a.batch * a.steps)
    ab = algorithmic_bytes(det, a.rows, a.cols, hits_per_frame)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel group by device time (CUDA events recorded between stages on the engine's stream)
    kstages = {k: v for k, v in stages.items() if k in ("color", "real", "chan", "pad", "cascade")}
    dom = max(kstages, key=kstages.get) if kstages else None
    alg = {"color": ab["in_u8"] + 4 * a.rows * a.cols * (3 if opts["colorSpace"] == "luv" else 1),
           "real": None, "chan": ab["chan_f32"], "pad": 0, "cascade": ab["B_det"]}
    # real-scale group: reads each real scale's source image once, writes the smoothed images later octaves resample from
    # and the real-scale channels; computed from the plan.  Its intermediate planes (M, O, U) are traffic,
    # not algorithmic bytes.
    reals = [s for s in info if s.is_real]
    np_img = 3 if opts["colorSpace"] == "luv" else 1
    px = [int(round(a.rows * s.scale / 4) * 4) * int(round(a.cols * s.scale / 4) * 4) for s in reals]
    alg["real"] = sum(4 * np_img * p for p in px) + sum(4 * np_img * p for p in px[:2]) + sum(4 * s.nchn * (s.w * s.h) for s in reals)
    roof = None
    traffic = None
    try:  # DRAM bytes of the dominant kernel group from the committed ncu --set full capture of this workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        wl = tr["workload"]
        if (wl["rows"], wl["cols"], wl["model"], wl["batch"], wl.get("operating_point", "fast")) == (a.rows, a.cols, a.model, a.batch, a.operating_point):
            members = {"color": ["k_color"], "real": REAL_GROUP, "chan": ["k_chan"], "cascade": ["k_cascade"]}.get(dom, [])
            traffic = sum(tr["per_step"][k]["dram_bytes"] for k in members if k in tr["per_step"]) or None
    except Exception:
        traffic = None
    if dom:
        achieved = alg[dom] * a.batch / (kstages[dom] / 1000.0) / 1e9
        roof = {"kernel": {"color": "k_color", "real": "real-scale group per octave: " + " + ".join(REAL_GROUP), "chan": "k_chan", "pad": "k_pad", "cascade": "k_cascade"}[dom],
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "dram read+write bytes of the kernel group per step (all its launches), ncu --set full, profiles/r1_traffic.json",
                "algorithmic_bytes_per_step": alg[dom] * a.batch,
                "peak_source": peak_src, "algorithmic_bytes_per_frame": alg[dom], "ms_per_launch_group": kstages[dom],
                "share_of_step": kstages[dom] / max(1e-9, sum(kstages.values())),
                "path_frac": ab["B_frame"] * fps / world / 1e9 / peak,
                "stage_ms": stages}
    cpu = None
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nf = a.cpu_frames or max(16, 8 * cores)  # ~10-30 s of CPU work
        nf = min(nf, 256)
        cf = [base[i % a.distinct] for i in range(nf)]
        dt, k, kind, _ = cpu_run(opts, clf, cf, cores)
        cpu = {"value": k / dt, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": f"{nf} of the same 1080p frames, frame-parallel over {cores} host threads, {dt:.1f} s"}
    line = {"metric": "frames_per_sec_1080p", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": 1000 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config, parallelism=f"batch-sharded x{world}", global_batch=a.batch * world),
            "mwindows_per_sec": fps * windows_per_frame / 1e6, "windows_per_frame": windows_per_frame,
            "trees_per_window": trees / max(1, windows), "hits_per_frame": hits_per_frame,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": a.batch * a.rows * a.cols * bpp,
                    "d2h_bytes_per_step": int(a.batch * 4 + 16 + 16 * hits_e2e / a.steps), "ms_per_step": 1000 * t_e2e / a.steps,
                    "stage_ms": stages_e2e},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "algorithmic_bytes": ab}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
