timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "short_atoms or known_answer or synthetic or dedup" > gpurun_out/r02l_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r02l_gputests.log | cut -c1-300
timeout 600 python bench.py --size-gib 1 --no-e2e --no-cpu --short-atoms > gpurun_out/r02l_bench_short.json 2> gpurun_out/r02l_bench_short.err; echo "rc=$?"; tail -3 gpurun_out/r02l_bench_short.err
YRB_FILTER_NO_SHORT=1 YR_B200_CACHE=0 timeout 600 python bench.py --size-gib 1 --no-e2e --no-cpu --no-parity --short-atoms > gpurun_out/r02l_bench_short_walk.json 2> gpurun_out/r02l_bench_short_walk.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench_short","bench_short_walk"):
    try:
      for l in open("gpurun_out/r02l_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["config"]["ms_breakdown"], d["config"].get("parity_checked"), d["config"]["setup"].get("prefilter"), d["config"]["scan_kernel"], d["config"]["matches"])
    except Exception as e: print(f, "ERR", e)
PY
