"""Experiment: two engines on one device, batches alternating between them (each has its own buffers and stream), to see how
much of the step's wave-quantisation / ramp-down holes the other stream's kernels fill.  python tools/two_engines.py [steps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import acf_b200, bench
from acf_b200 import synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
opts = bench.model_opts("face80"); clf = bench.make_clf(opts, "face80", 2048, "hits")
dets = []
for i in range(ne):
    d = acf_b200.Detector(acf_b200.Model.create(opts, clf), device=0, max_rows=1080, max_cols=1920, max_batch=batch)
    d.setHitCapacity(8192); d.setDoNonMaximaSuppression(True); dets.append(d)
base = synth.frames("shapes", 16, 1080, 1920, seed0=100)
host = np.stack([base[i % 16] for i in range(batch)])
dev = torch.from_numpy(host).cuda()
for d in dets:
    for _ in range(3):
        d.submit(dev.data_ptr(), batch, 1080, 1920, True); d.collect_arrays(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
q = []
for k in range(steps):
    d = dets[k % ne]
    d.submit(dev.data_ptr(), batch, 1080, 1920, True); q.append(d)
    if len(q) > ne:
        q.pop(0).collect_arrays(batch)
while q:
    q.pop(0).collect_arrays(batch)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"engines {ne} batch {batch}: {steps * batch / dt:.0f} frames/s, {1000 * dt / steps:.2f} ms per batch")
