timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "rc=$?"
timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02o_bench_b.json 2> gpurun_out/r02o_bench_b.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench","bench_b"):
    for l in open("gpurun_out/r02o_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["config"]["ms_breakdown"])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or dedup or sharded" 2>&1 | tail -2
