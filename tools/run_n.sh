#!/bin/bash
# usage: run_n.sh N tag [env...] [-- extra bench args]
N=$1; TAG=$2; shift 2
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity $EXTRA > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
python - "$TAG" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], round(d["value"]), round(d["ms_per_step"], 2), d["ms_per_step_per_rank"], "e2e", round(d["e2e"]["value"]))
except Exception as ex:
    print(sys.argv[1], "failed", ex)
PY
