#!/bin/bash
# The driver's scaling run for one N (under gpurun --gpus N): bench.py at N GPUs exactly as the driver launches it.
# usage: bash scripts/r02_scale.sh N [tag]
N=${1:-8}; TAG=${2:-r02}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
fi
echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("N=$N value %.1f GB/s  ms/step %.3f  kernel frac %.3f  e2e %s  weak %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value"), (d.get("weak") or {}).get("value")))
        print(d["config"]["ms_breakdown"], d["config"].get("parity_checked"), d["config"].get("flow"))
PY
tail -5 gpurun_out/${TAG}_bench_n$N.err
