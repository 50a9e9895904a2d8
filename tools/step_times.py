"""Per-batch completion times of the resident submit / collect loop (diagnostic): python tools/step_times.py [steps] [depth]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import acf_b200, bench
from acf_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
batch = 256
opts = bench.model_opts("face80"); clf = bench.make_clf(opts, "face80", 2048, "hits")
d = acf_b200.Detector(acf_b200.Model.create(opts, clf), device=0, max_rows=1080, max_cols=1920, max_batch=batch)
d.setHitCapacity(8192); d.setDoNonMaximaSuppression(True)
base = synth.frames("shapes", 16, 1080, 1920, seed0=100)
dev = torch.from_numpy(np.stack([base[i % 16] for i in range(batch)])).cuda()
for _ in range(4):
    d.submit(dev.data_ptr(), batch, 1080, 1920, True); d.collect_arrays(batch)
for rep in range(3):
    torch.cuda.synchronize()
    ts, sub = [], []
    t0 = time.perf_counter()
    for k in range(min(depth, steps)):
        a = time.perf_counter(); d.submit(dev.data_ptr(), batch, 1080, 1920, True); sub.append(time.perf_counter() - a)
    for k in range(steps):
        if k + depth < steps:
            a = time.perf_counter(); d.submit(dev.data_ptr(), batch, 1080, 1920, True); sub.append(time.perf_counter() - a)
        d.collect_arrays(batch)
        ts.append(time.perf_counter() - t0)
    dt = np.diff([0.0] + ts) * 1000
    print(f"rep {rep}: {1000 * ts[-1] / steps:.2f} ms/step; per-step min {dt.min():.1f} median {np.median(dt):.1f} max {dt.max():.1f}; "
          f"submit call ms: median {1000 * np.median(sub):.2f} max {1000 * max(sub):.2f}; steps > 25 ms: {[int(i) for i in np.nonzero(dt > 25)[0]][:12]}")
