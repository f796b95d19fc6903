#!/bin/bash
# Round-2 kernel sweep on one B200: parity of the new filter-kernel variants, then scan GB/s per variant.
# usage (under gpurun): bash scripts/r02_kernel_sweep.sh "<variants>" [size_gib]
set -u
VARS=${1:-"0 3 6"}
SZ=${2:-4}
mkdir -p gpurun_out
for v in $VARS; do
  echo "=== variant $v: parity" | tee -a gpurun_out/sweep.log
  YRB_FILTER_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/sweep.log
  echo "=== variant $v: bench ${SZ} GiB" | tee -a gpurun_out/sweep.log
  YRB_FILTER_VARIANT=$v timeout 300 python bench.py --size-gib $SZ --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/sweep_v$v.json 2> gpurun_out/sweep_v$v.err
  python - <<PY | tee -a gpurun_out/sweep.log
import json
try:
    l=json.loads(open("gpurun_out/sweep_v$v.json").read().strip().splitlines()[-1])
    print("variant $v: value %.1f GB/s  scan kernel %.1f GB/s frac %.4f  breakdown %s matches %s clocks %s" % (l["value"], l["roofline"]["achieved"], l["roofline"]["frac"], l["config"]["ms_breakdown"], l["config"]["matches_rank0"], l["clocks"]))
except Exception as e:
    print("variant $v: no result", e); print(open("gpurun_out/sweep_v$v.err").read()[-2000:])
PY
done
