#!/bin/bash
# tools/sweep_env.sh "VAR=a VAR2=b" "VAR=c" ... : one short resident + e2e bench per environment string (GPU box)
for cfg in "$@"; do
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --no-parity > gpurun_out/_sw.json 2> gpurun_out/_sw.err || tail -c 300 gpurun_out/_sw.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/_sw.json"))
    h = d.get("host_ms_per_step", {})
    print(f"{sys.argv[1]:60s} fps {d['value']:8.0f}  ms {d['ms_per_step']:6.2f} (events {d.get('ms_per_step_device_events', 0):6.2f})  e2e {d['e2e']['value']:8.0f}  stages "
          + " ".join(f"{k}={v:.2f}" for k, v in d['roofline']['stage_ms'].items()) + "  host " + " ".join(f"{v:.2f}" for v in h.values()))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
