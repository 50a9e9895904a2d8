#!/usr/bin/env python
"""bench.py -- 1080p frames/s and Mwindows/s of the chnsPyramid + acfDetect path on B200.

A "step" is one pass of the hot path over one batch of synthetic frames (BASELINE.json configs[1]:
1080p, batch 256 per GPU, FACE80-shaped 7-channel model, full 31-scale pyramid + cascade + NMS).
  value : frames/s with the u8 frames already resident in HBM (acfb_submit on_device=1 + acfb_collect)
  e2e   : the same through the public C ABI with HOST (pinned) frames -- H2D of every frame and D2H of
          the hit lists inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"
  parity_checked : frames of the LAST timed batch whose boxes and scores were compared (bit for bit) with the
          CPU oracle after the timed region
  other_configs : BASELINE.json configs[2] (4K x64) and configs[3] (INRIA-shaped 10-channel model, 1080p x256)
          measured the same way in the same run (N = 1 only)
Multi-GPU (torchrun, one rank per GPU): frames shard by batch (weak scaling, 256 per GPU), no data-path
collective; the only exchange is the gather of the per-frame detection lists on rank 0.
--impl reference times the reference's own CPU implementation (oracle/_ref, native SSE arithmetic)
on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REAL_GROUP = ["k_resample", "k_resample_x", "k_resample_y", "k_front", "k_smooth", "k_gradmag", "k_trix", "k_triyhist_tma", "k_triyhist", "k_hist"]
GROUP_KERNELS = {"color": ["k_color"], "real": REAL_GROUP, "chan": ["k_chan", "k_pad"], "cascade": ["k_cascade_tile", "k_cascade_tail_win", "k_cascade_tail", "k_cascade", "k_post"]}


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
                    help="synthetic cascade: hits (headline: ~12 trees/window, the survivors of 54 rejector trees walk all 2048 trees and "
                         "become ~200 raw hits per frame), fast (the same rejectors, no hits) or deep (~75 trees/window, no hits)")
    ap.add_argument("--input-format", default="rgb", choices=["rgb", "gray"],
                    help="frames handed to the detector: RGB24 (default, the headline) or GRAY8 (one third of the PCIe bytes; "
                         "gray / orig models only, chnsPyramid.cpp:234-244)")
    ap.add_argument("--e2e-format", default="nv12", choices=["nv12", "same"],
                    help="host frames of the end-to-end measurement: NV12 (default; what cameras / decoders deliver, 1.5 bytes per pixel over "
                         "PCIe, converted on the device like cv::cvtColor(COLOR_YUV2RGB_NV12)) or the resident format (--input-format); "
                         "the RGB24 figure is reported beside it as e2e_rgb24")
    ap.add_argument("--no-nms", action="store_true", help="return the raw hits (default: bbNms + prune as acf-detect runs the detector)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs[2] / configs[3]")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic: every rank keeps its boxes (no exchange at all)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU baseline sample (0 = 2 per core)")
    return ap.parse_args()


def model_opts(name):
    from acf_b200 import synth
    return {"face80": lambda: synth.face_opts(80), "face80c": lambda: synth.face_opts(80, True), "inria": synth.inria_opts}[name]()


def make_clf(opts, model, trees, point):
    from acf_b200 import synth
    if point == "deep":
        return synth.make_classifier(opts, trees, 2, seed=1, drift=-0.113, gain=0.23)
    if model == "inria":  # the INRIA-shaped stand-in has its own calibrated rejector count (48): survivors become hits
        return synth.make_classifier(opts, trees, 2, seed=1, n_reject=trees if point == "fast" else None)
    return synth.make_classifier(opts, trees, 2, seed=1, n_reject=54 if point == "hits" else None)


def workload_config(a, model, rows, cols, batch, opts, world):
    from acf_b200 import synth
    bpp = 3 if a.input_format == "rgb" else 1
    return {"workload": f"{rows}x{cols} synthetic 'shapes' frames, batch {batch}/GPU, {model} "
                        f"({synth.n_channels(opts)} channels, {a.trees} depth-2 trees), full pyramid + cascade"
                        + ("" if a.no_nms else " + bbNms / prune"),
            "frames_per_step_per_gpu": batch, "distinct_frames": a.distinct, "model": model, "operating_point": a.operating_point,
            "nms": not a.no_nms,
            "resident_frame_format": "RGB24" if a.input_format == "rgb" else "GRAY8", "e2e_host_frame_format": "NV12" if (a.e2e_format == "nv12" and a.input_format == "rgb") else "as resident",
            "l2_policy": f"inputs larger than L2 ({batch * rows * cols * bpp / 1e9:.2f} GB of u8 frames per step per GPU)",
            "parallelism": f"batch-sharded x{world}", "global_batch": batch * world,
            "engine_pipelines": int({"1": 1, "2": 2}.get(os.environ.get("ACFB_PIPELINES", ""), 3)), "batches_in_flight": int({"1": 3, "2": 4}.get(os.environ.get("ACFB_PIPELINES", ""), 6)),
            "detection_gather": ("none (one GPU)" if world == 1 else
                                 ("engine (acfb_dist_collect): " + ("shared-memory ring, single node" if os.environ.get("ACFB_DIST_EXCHANGE") == "shm" else "ncclAllGather of k_post's device records")
                                  if not a.no_nms else "torch.distributed gather of host lists (raw hits)"))}


def algorithmic_bytes(det, rows, cols, hits_per_frame):
    """B_pyr = in_u8 + chan_f32 ; B_det = chan_f32 + 24*hits ; chan_f32 = dense padded channel bytes."""
    info, _ = det.plan(rows, cols)
    chan = sum(s.nchn * s.w * s.h for s in info) * 4
    in_u8 = rows * cols * 3
    return dict(in_u8=in_u8, chan_f32=chan, B_pyr=in_u8 + chan, B_det=chan + 24 * hits_per_frame,
                B_frame=in_u8 + 2 * chan + 24 * hits_per_frame)


class ClockSampler:
    def __init__(self, index):
        self.proc = None
        self.index = index

    def start(self):
        if os.environ.get("BENCH_NO_CLOCKS") == "1":
            return
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
        sm_sorted = sorted(sm)  # keep the samples taken under load (upper half) for the median
        med = float(np.median(sm_sorted[len(sm_sorted) // 2:])) if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_run(opts, clf, frames, threads, repeats=1, frames_1thread=0, courtesy=False):
    """The reference's CPU path timed by oracle/cpu_bench.cpp: a C++ std::thread driver over the reference's own toolbox objects
    (oracle/_ref, native SSE arithmetic), frame-parallel like src/app/acf/acf.cpp:443-455 -- no Python in the timed loop.
    Returns the driver's JSON (fps, fps_1thread, stage_ms, ...) + "kind": "reference" | "port"."""
    import ctypes
    import tempfile
    from oracle import oracle as O
    exe = os.path.join(ROOT, "oracle", "_ref", "cpu_bench_native_o3" if courtesy else "cpu_bench_native")
    kind = "reference"
    if not os.path.exists(exe):
        exe, kind = os.path.join(ROOT, "oracle", "cpu_bench_port"), "port"
        if not os.path.exists(exe):
            O.build(ref=False)
    oo = O.opts_from_dict(opts)
    fids = np.ascontiguousarray(clf["fids"], np.uint32)
    with tempfile.NamedTemporaryFile(suffix=".acfb", delete=False) as f:
        f.write(b"ACFB" + np.uint32(1).tobytes() + bytes(oo))
        f.write(np.array([fids.shape[0], fids.shape[1], int(clf["treeDepth"])], np.int32).tobytes())
        f.write(fids.tobytes()); f.write(np.ascontiguousarray(clf["thrs"], np.float32).tobytes())
        f.write(np.ascontiguousarray(clf["child"], np.uint32).tobytes()); f.write(np.ascontiguousarray(clf["hs"], np.float32).tobytes())
        fr = np.ascontiguousarray(np.stack(frames), np.uint8)
        f.write(np.array([fr.shape[0], fr.shape[1], fr.shape[2]], np.int32).tobytes())
        f.write(fr.tobytes())
        path = f.name
    try:
        out = subprocess.run([exe, path, str(threads), str(repeats), str(frames_1thread)], check=True, capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    res = json.loads(out.strip().splitlines()[-1])
    res["kind"] = kind
    res["exe"] = os.path.relpath(exe, ROOT)
    return res


def oracle_detections(opts, clf, frame, nms, max_det=10):
    """the CPU oracle's answer for one frame: (rects, scores) as Detector::operator() returns them"""
    from oracle import oracle as O
    orc = O.Oracle("port")
    P = orc.pyramid(opts, frame)
    dets, _, _, total = P.detect(clf, cap=1 << 20)
    P.close()
    if nms and dets:
        dets = orc.prune(orc.nms(dets, overlap=opts["nms_overlap"], greedy=opts["nms_type"] == "maxg", ovr_union=opts["nms_ovrDnm"] != "min"), max_det, 0.0)
    return [tuple(d[:4]) for d in dets], np.array([d[4] for d in dets], np.float32), total


# ------------------------------------------------------------------------------------------ one workload on this rank's GPU
class Workload:
    def __init__(self, a, model, rows, cols, batch, rank, local, dist):
        import torch
        import acf_b200
        from acf_b200 import synth
        self.a, self.model, self.rows, self.cols, self.batch, self.rank, self.dist = a, model, rows, cols, batch, rank, dist
        self.torch = torch
        self.opts = model_opts(model)
        self.clf = make_clf(self.opts, model, a.trees, a.operating_point)
        self.det = acf_b200.Detector(acf_b200.Model.create(self.opts, self.clf), device=local, max_rows=rows, max_cols=cols, max_batch=batch)
        self.det.setHitCapacity(8192)
        if not a.no_nms:
            self.det.setDoNonMaximaSuppression(True)
        # synthetic frames: `distinct` seeded frames tiled to the batch; different seeds per rank
        self.base = synth.frames("shapes", a.distinct, rows, cols, seed0=100 + 1000 * rank)
        self.bpp = 3 if a.input_format == "rgb" else 1
        if self.bpp == 1:
            self.det.setInputFormat("gray")
        self.host = torch.empty((batch, rows, cols, self.bpp), dtype=torch.uint8).pin_memory()
        hv = self.host.numpy()
        for i in range(batch):
            hv[i] = self.base[i % a.distinct] if self.bpp == 3 else self.base[i % a.distinct][:, :, 1:2]
        self.dev = self.host.cuda(non_blocking=False)
        self.fmt = "rgb" if self.bpp == 3 else "gray"
        self.host_nv12 = None
        if self.bpp == 3 and a.e2e_format == "nv12" and rows % 2 == 0 and cols % 2 == 0:
            self.base_nv12 = [synth.rgb_to_nv12(f) for f in self.base]
            self.host_nv12 = torch.empty((batch, rows * 3 // 2, cols), dtype=torch.uint8).pin_memory()
            nv = self.host_nv12.numpy()
            for i in range(batch):
                nv[i] = self.base_nv12[i % a.distinct]
        # multi-GPU: the boxes are gathered by the engine itself (acfb_dist_*: ncclAllGather of the device buffer k_post wrote, enqueued
        # at submit time on a communication stream); torch.distributed only carries the 128-byte NCCL unique id
        self.engine_gather = False
        if dist is not None and not a.no_nms and not a.no_gather:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(acf_b200.Detector.dist_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            self.det.dist_init_rank(uid.cpu().numpy().tobytes(), rank, dist.get_world_size())
            self.engine_gather = True
        self.stream = torch.cuda.ExternalStream(self.det.stream(), device=local)
        self.cap = 1 << 19
        self.last = None  # (dets, counts) of the most recent collected batch
        self.hostt = {"wait_ms": 0.0, "tail_ms": 0.0, "collect_ms": 0.0, "n": 0}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def gather(self, dets, counts):
        """gather of the variable-length detection lists on rank 0 (counts, then records)"""
        if self.dist is None or self.a.no_gather:
            return
        from acf_b200 import dist as adist
        adist.gather_detection_arrays(dets, counts, self.dist, "cuda", frame0=self.rank * self.batch)

    def collect(self):
        t0 = time.perf_counter()
        if self.engine_gather:
            # every rank: its own collect + the gather; rank 0 receives all ranks' boxes in global frame order (its own come first)
            dets, counts, total = self.det.dist_collect_arrays(self.batch, cap=self.cap)
            w, t = self.det.collect_times()
            self.hostt["wait_ms"] += w; self.hostt["tail_ms"] += t
            if self.rank == 0:
                own = int(counts[:self.batch].sum())
                self.last = (dets[:own].copy(), counts[:self.batch].copy())
                self.gathered_boxes = total
                total = own
        else:
            dets, counts, total = self.det.collect_arrays(self.batch, cap=self.cap)
            w, t = self.det.collect_times()
            self.hostt["wait_ms"] += w; self.hostt["tail_ms"] += t
            self.last = (dets.copy(), counts.copy())
            self.gather(dets, counts)
        self.hostt["collect_ms"] += 1000 * (time.perf_counter() - t0); self.hostt["n"] += 1
        return total

    def timed(self, on_device, steps, stage_timing=False, nv12=False):
        """K steps through the public asynchronous API with up to three batches in flight: while the host orders / rescales the
        hits of step k, the kernels of step k+1 already run; with host frames the H2D copy of step k+1 (copy stream) overlaps the
        kernels of step k -- every step still copies its own frames from pinned host memory inside the timed region."""
        torch, det = self.torch, self.det
        self.hostt = {"wait_ms": 0.0, "tail_ms": 0.0, "collect_ms": 0.0, "n": 0}
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = det.launch_count()
        t0 = time.perf_counter()
        e0.record(self.stream)
        tot = 0
        stage_acc = {}
        ptr = self.dev.data_ptr() if on_device else (self.host_nv12.data_ptr() if nv12 else self.host.data_ptr())
        det.setInputFormat("nv12" if nv12 else self.fmt)
        self.last_fmt = "nv12" if nv12 else self.fmt
        if stage_timing:  # instrumented: one batch at a time, CUDA events between the kernel groups on the engine's stream
            for _ in range(steps):
                det.submit(ptr, self.batch, self.rows, self.cols, on_device)
                tot += self.collect()
                for nme, ms in det.stage_times():
                    stage_acc[nme] = stage_acc.get(nme, 0.0) + ms
        else:
            depth = int(os.environ.get("BENCH_DEPTH", {"1": "2", "2": "3"}.get(os.environ.get("ACFB_PIPELINES", ""), "5")))  # batches submitted ahead of the one being collected (the engine keeps two per pipeline in flight)
            for k in range(min(depth, steps)):
                det.submit(ptr, self.batch, self.rows, self.cols, on_device)
            for k in range(steps):
                if k + depth < steps:
                    det.submit(ptr, self.batch, self.rows, self.cols, on_device)
                tot += self.collect()
        e1.record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the stream is idle while the host orders / rescales hits, so the larger of device-event span and wall clock covers the step
        t = torch.tensor([ms, wall * 1000.0], device="cuda", dtype=torch.float64)
        self.per_rank_ms = [max(ms, wall * 1000.0) / steps]
        if self.dist is not None:
            allt = [torch.zeros_like(t) for _ in range(self.dist.get_world_size())]
            self.dist.all_gather(allt, t)
            self.per_rank_ms = [round(float(max(x[0].item(), x[1].item())) / steps, 3) for x in allt]
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        hn = max(1, self.hostt["n"])
        return dict(ms=max(t[0].item(), t[1].item()), ms_events=t[0].item(), ms_wall=t[1].item(), dets=tot, launches=det.launch_count() - l0,
                    stages={k: v / steps for k, v in stage_acc.items()},
                    host_ms_per_step={"blocked_on_device": self.hostt["wait_ms"] / hn, "host_tail (order, rescale, nms)": self.hostt["tail_ms"] / hn,
                                      "collect_call_incl_python_and_gather": self.hostt["collect_ms"] / hn})

    def run(self, steps, warmup, stages=True):
        det = self.det
        for _ in range(max(3, warmup)):
            det.submit(self.dev.data_ptr(), self.batch, self.rows, self.cols, True)
            self.collect()  # also warms up the NCCL communicator (lazy initialisation would land in the timed region)
        # untimed: the same pipelined submit / collect pattern as the timed region, long enough to touch every slot of both pipelines
        # (the second pipeline and each slot's buffers are created at first use: ~24 GB of cudaMalloc + clears must not land in the timed steps)
        self.timed(True, 12)
        dev = self.timed(True, steps)
        dev["per_rank_ms"] = list(self.per_rank_ms)
        raw_hits, trees, windows = det.last_hit_count()
        st = {}
        if stages:
            # per-kernel-group device times for the roofline: a second, instrumented pass over the same K steps (instrumentation
            # serialises the streams the engine otherwise overlaps, so these durations are per kernel group, not per step)
            det.enable_stage_timing(True)
            self.timed(True, 1, True)
            st = self.timed(True, steps, True)["stages"]
            det.enable_stage_timing(False)
        self.timed(False, 12)  # untimed: batches in flight on both pipelines, so every slot's staging buffer exists before the timed region
        e2e = self.timed(False, steps)
        e2e_nv12 = None
        if self.host_nv12 is not None:  # last, so that parity() checks frames of this pass
            self.timed(False, 12, nv12=True)
            e2e_nv12 = self.timed(False, steps, nv12=True)
        return dict(dev=dev, e2e=e2e, e2e_nv12=e2e_nv12, stages=st, raw_hits_per_frame=raw_hits / self.batch, trees=trees, windows=windows)

    def parity(self, n_frames=8):
        """frames of the last collected batch (the e2e pass's final step) against the CPU oracle, outside any timed region"""
        dets, counts = self.last
        offs = np.concatenate([[0], np.cumsum(counts)])
        stride = max(1, self.batch // n_frames)
        checked, bad, hits = 0, [], 0
        for k in range(min(n_frames, self.batch)):
            f = min(self.batch - 1, k * stride + (k % stride))
            frame = self.base[f % self.a.distinct]
            if self.last_fmt == "nv12":  # the definition of the device conversion: cv::cvtColor(COLOR_YUV2RGB_NV12)'s integer formula
                from acf_b200 import synth
                frame = synth.nv12_to_rgb(self.base_nv12[f % self.a.distinct])
            if self.bpp == 1:
                frame = np.repeat(frame[:, :, 1:2], 3, axis=2)
            rects, scores, total = oracle_detections(self.opts, self.clf, frame, not self.a.no_nms)
            d = dets[offs[f]:offs[f + 1]]
            mine = [(int(x), int(y), int(w), int(h)) for x, y, w, h in zip(d["x"], d["y"], d["w"], d["h"])]
            ok = (sorted(mine) == sorted(rects)) if not self.a.no_nms else (mine == rects)
            ok = ok and np.array_equal(np.sort(d["score"]), np.sort(scores))
            checked += 1; hits += total
            if not ok:
                bad.append(int(f))
        return {"frames": checked, "mismatched_frames": bad, "oracle_raw_hits_in_sample": int(hits), "ok": not bad, "input_format": self.last_fmt,
                "what": "boxes and scores of sampled frames of the last timed (e2e) batch == CPU oracle (port of the reference's exact-math build), bit for bit"}


def roofline_of(ab, stages, batch, fps, world, peak, peak_src, traffic_json, wl_key):
    """roofline of the dominant kernel group + the SURVEY 8(d) figures (pyramid producer on B_pyr, cascade on B_det, path on B_frame)"""
    kst = {k: v for k, v in stages.items() if k in ("color", "real", "chan", "pad", "cascade")}
    if not kst:
        return None
    t_pyr = sum(v for k, v in kst.items() if k != "cascade") / 1000.0
    t_det = kst.get("cascade", 0.0) / 1000.0
    dom = "cascade" if t_det >= t_pyr else "pyramid"
    alg = ab["B_det"] if dom == "cascade" else ab["B_pyr"]
    t_dom = t_det if dom == "cascade" else t_pyr
    traffic = None
    try:  # DRAM bytes of the dominant group from the committed ncu --set full capture of this workload
        tr = json.load(open(traffic_json))
        wl = tr["workload"]
        if (wl["rows"], wl["cols"], wl["model"], wl["batch"], wl.get("operating_point")) == wl_key:
            members = GROUP_KERNELS["cascade"] if dom == "cascade" else GROUP_KERNELS["color"] + GROUP_KERNELS["real"] + GROUP_KERNELS["chan"]
            traffic = sum(tr["per_step"][k]["dram_bytes"] for k in members if k in tr["per_step"]) or None
    except Exception:
        traffic = None
    achieved = alg * batch / max(t_dom, 1e-9) / 1e9
    return {"kernel": "k_cascade_tile + k_cascade_tail_win (acfDetect1)" if dom == "cascade" else "pyramid producer: k_color + real-scale group + k_chan (chnsPyramid)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_note": "dram read+write bytes of the group's launches in one step (ncu, dram__bytes_read.sum + dram__bytes_write.sum per launch), " + os.path.relpath(traffic_json, ROOT),
            "algorithmic_bytes_per_step": alg * batch, "algorithmic_bytes_per_frame": alg, "bytes_definition": "SURVEY 8(d): B_pyr = in_u8 + chan_f32, B_det = chan_f32 + 24 hits",
            "ms_per_launch_group": t_dom * 1000.0, "share_of_step": t_dom / max(1e-9, t_pyr + t_det),
            "peak_source": peak_src,
            "pyramid_frac": ab["B_pyr"] * batch / max(t_pyr, 1e-9) / 1e9 / peak, "cascade_frac": ab["B_det"] * batch / max(t_det, 1e-9) / 1e9 / peak,
            "path_frac": ab["B_frame"] * fps / world / 1e9 / peak, "stage_ms": stages}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    opts = model_opts(a.model)
    config = workload_config(a, a.model, a.rows, a.cols, a.batch, opts, world)

    if a.impl == "reference":
        if rank != 0:
            return
        from acf_b200 import synth
        clf = make_clf(opts, a.model, a.trees, a.operating_point)
        cores = os.cpu_count() or 1
        nf = a.cpu_frames or max(16, 4 * cores)
        frames = synth.frames("shapes", min(nf, a.distinct), a.rows, a.cols, seed0=100)
        frames = [frames[i % len(frames)] for i in range(nf)]
        r = cpu_run(opts, clf, frames, cores, repeats=a.steps, frames_1thread=0)  # the driver warms up once by itself
        fps, kind = r["fps"], r["kind"]
        line = {"impl": "reference", "metric": "frames_per_sec_1080p", "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * r["seconds"] / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "threads": cores, "kind": kind,
                                 "sample": f"{nf} frames per step, frame-parallel over {cores} std::threads ({r['exe']}), native SSE arithmetic, "
                                           "pyramid + cascade (bbNms of a few hundred boxes is negligible)",
                                 "stage_ms": r["stage_ms"]},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic_json = os.path.join(ROOT, "profiles", "r2_traffic.json")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # samples clocks / throttle reasons through warm-up and every timed region
    W = Workload(a, a.model, a.rows, a.cols, a.batch, rank, local, dist)
    R = W.run(a.steps, a.warmup)
    parity = None
    if rank == 0 and not a.no_parity:
        parity = W.parity(8)
    info_windows = R["windows"] / a.batch
    frames_total = a.batch * world * a.steps
    fps = frames_total / (R["dev"]["ms"] / 1000.0)
    E = R["e2e_nv12"] or R["e2e"]  # the headline end-to-end figure: NV12 host frames when available
    fps_e2e = frames_total / (E["ms"] / 1000.0)
    fps_e2e_rgb = frames_total / (R["e2e"]["ms"] / 1000.0)
    ab = algorithmic_bytes(W.det, a.rows, a.cols, R["raw_hits_per_frame"])
    roof = roofline_of(ab, R["stages"], a.batch, fps, world, peak, peak_src, traffic_json, (a.rows, a.cols, a.model, a.batch, a.operating_point))
    h2d_rgb = a.batch * a.rows * a.cols * W.bpp
    h2d = a.batch * a.rows * a.cols * 3 // 2 if R["e2e_nv12"] else h2d_rgb
    e2e_fmt = "NV12" if R["e2e_nv12"] else ("RGB24" if W.bpp == 3 else "GRAY8")
    opts0, clf0, base0, distinct = W.opts, W.clf, W.base, a.distinct
    dets_per_step = E["dets"] / a.steps
    launches = R["dev"]["launches"]
    del W
    torch.cuda.empty_cache()

    other = None
    if world == 1 and not a.no_other_configs and (a.rows, a.cols, a.model, a.batch) == (1080, 1920, "face80", 256):
        other = {}
        for name, model, rows, cols, batch in (("cfg3", "face80", 2160, 3840, 64), ("cfg4", "inria", 1080, 1920, 256)):
            try:
                Wo = Workload(a, model, rows, cols, batch, rank, local, None)
                k = max(3, a.steps // 2)
                Ro = Wo.run(k, 3)
                f = batch * k / (Ro["dev"]["ms"] / 1000.0)
                fe = batch * k / ((Ro["e2e_nv12"] or Ro["e2e"])["ms"] / 1000.0)
                abo = algorithmic_bytes(Wo.det, rows, cols, Ro["raw_hits_per_frame"])
                ro = roofline_of(abo, Ro["stages"], batch, f, 1, peak, peak_src, traffic_json, None)
                other[name] = {"workload": workload_config(a, model, rows, cols, batch, Wo.opts, 1)["workload"], "steps": k,
                               "value": f, "unit": "frames/s", "ms_per_step": Ro["dev"]["ms"] / k, "e2e": fe, "e2e_rgb24": batch * k / (Ro["e2e"]["ms"] / 1000.0), "windows_per_frame": Ro["windows"] / batch,
                               "mwindows_per_sec": f * Ro["windows"] / batch / 1e6, "trees_per_window": Ro["trees"] / max(1, Ro["windows"]),
                               "hits_per_frame": Ro["raw_hits_per_frame"], "path_frac": ro["path_frac"] if ro else None,
                               "pyramid_frac": ro["pyramid_frac"] if ro else None, "cascade_frac": ro["cascade_frac"] if ro else None,
                               "stage_ms": Ro["stages"], "algorithmic_bytes": abo,
                               "parity_checked": None if a.no_parity else Wo.parity(2)}
                del Wo
                torch.cuda.empty_cache()
            except Exception as ex:  # a failed side workload must not hide the headline
                other[name] = {"error": repr(ex)}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nf = a.cpu_frames or max(16, 8 * cores)  # ~10-30 s of CPU work
        nf = min(nf, 256)
        cf = [base0[i % distinct] for i in range(nf)]
        r = cpu_run(opts0, clf0, cf, cores, repeats=1, frames_1thread=min(4, nf))
        cpu = {"value": r["fps"], "unit": "frames/s", "cores": cores, "threads": cores, "kind": r["kind"],
               "sample": f"{nf} of the same 1080p frames, frame-parallel over {cores} std::threads ({r['exe']}: the reference's toolbox objects, -O2, "
                         f"baseline x86-64 like the reference's own build), {r['seconds']:.1f} s",
               "fps_per_thread": r["fps_per_thread"], "fps_1thread": r.get("fps_1thread"), "stage_ms": r["stage_ms"]}
        try:  # courtesy row: the same objects with -O3 -mavx2
            r3 = cpu_run(opts0, clf0, cf[:max(cores, nf // 2)], cores, repeats=1, courtesy=True)
            cpu["courtesy_o3_avx2"] = {"value": r3["fps"], "exe": r3["exe"]}
        except Exception as ex:
            cpu["courtesy_o3_avx2"] = {"error": repr(ex)[:200]}
    line = {"metric": "frames_per_sec_1080p", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": R["dev"]["ms"] / a.steps, "ms_per_step_device_events": R["dev"]["ms_events"] / a.steps, "ms_per_step_per_rank": R["dev"].get("per_rank_ms"), "host_ms_per_step": R["dev"]["host_ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config,
            "mwindows_per_sec": fps * info_windows / 1e6, "windows_per_frame": info_windows,
            "trees_per_window": R["trees"] / max(1, R["windows"]), "hits_per_frame": R["raw_hits_per_frame"],
            "detections_per_frame": dets_per_step / a.batch,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "host_frame_format": e2e_fmt,
                    "d2h_bytes_per_step": int(a.batch * 4 + 16 + 16 * R["raw_hits_per_frame"] * a.batch), "ms_per_step": E["ms"] / a.steps,
                    "h2d_gbs_per_rank": h2d * a.steps / (E["ms"] / 1000.0) / 1e9, "host_ms_per_step": E["host_ms_per_step"]},
            "e2e_rgb24": {"value": fps_e2e_rgb, "unit": "frames/s", "h2d_bytes_per_step": h2d_rgb, "ms_per_step": R["e2e"]["ms"] / a.steps,
                          "h2d_gbs_per_rank": h2d_rgb * a.steps / (R["e2e"]["ms"] / 1000.0) / 1e9},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity_checked": parity,
            "other_configs": other, "algorithmic_bytes": ab}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
