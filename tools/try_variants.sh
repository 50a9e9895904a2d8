#!/bin/bash
# time library variants (acf_b200/libvar_*.so) with bench.py; prints step ms and per-stage ms
cp acf_b200/libacf_b200.so /tmp/lib_keep.so
for v in acf_b200/libvar_*.so; do
  cp $v acf_b200/libacf_b200.so
  for op in fast; do
    python bench.py --no-cpu-baseline --operating-point $op 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.load(sys.stdin); print('$v','$op', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:round(x,2) for k,x in d['roofline']['stage_ms'].items()})"
  done
done
cp /tmp/lib_keep.so acf_b200/libacf_b200.so
