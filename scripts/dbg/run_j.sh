YR_B200_CACHE=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02j_bench_v3.json 2> gpurun_out/r02j_bench_v3.err; echo "rc=$?"
YR_B200_CACHE=0 YRB_FILTER_DENSE_FILL=0.1 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02j_bench_v5.json 2> gpurun_out/r02j_bench_v5.err; echo "rc=$?"
YR_B200_CACHE=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02j_bench_v3b.json 2> gpurun_out/r02j_bench_v3b.err; echo "rc=$?"
YR_B200_CACHE=0 YRB_FILTER_DENSE_FILL=0.1 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02j_bench_v5b.json 2> gpurun_out/r02j_bench_v5b.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench_v3","bench_v5","bench_v3b","bench_v5b"):
    try:
      for l in open("gpurun_out/r02j_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["config"]["ms_breakdown"], d["config"]["setup"].get("prefilter",{}).get("variant"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 900 python scripts/cli_feeder.py 20000 600 16 > gpurun_out/r02j_cli_feeder.jsonl 2> gpurun_out/r02j_cli_feeder.err; echo "cli rc=$?"; cat gpurun_out/r02j_cli_feeder.jsonl | cut -c1-300
