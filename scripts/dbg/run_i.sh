timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02i_gputests.log
timeout 600 python scripts/fuzz_sweep.py 200 330 240 > gpurun_out/r02i_fuzz.log 2> gpurun_out/r02i_fuzz.err; echo "fuzz rc=$?"; tail -1 gpurun_out/r02i_fuzz.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "rc=$?"
YR_B200_CACHE=0 YRB_FILTER_VARIANT=5 timeout 600 python bench.py --no-e2e --no-cpu --no-parity --size-gib 4 > gpurun_out/r02i_bench_v5.json 2> gpurun_out/r02i_bench_v5.err; echo "rc=$?"
YR_B200_CACHE=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity --size-gib 4 > gpurun_out/r02i_bench_v3.json 2> gpurun_out/r02i_bench_v3.err; echo "rc=$?"
YRB_L2_PERSIST=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity --size-gib 4 > gpurun_out/r02i_bench_nol2.json 2> gpurun_out/r02i_bench_nol2.err; echo "rc=$?"
timeout 600 python bench.py --no-e2e --no-cpu --no-parity --size-gib 1 --kernel 1 > gpurun_out/r02i_bench_walk.json 2> gpurun_out/r02i_bench_walk.err; echo "rc=$?"
YRB_L2_PERSIST=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity --size-gib 1 --kernel 1 > gpurun_out/r02i_bench_walk_nol2.json 2> gpurun_out/r02i_bench_walk_nol2.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench","bench_v5","bench_v3","bench_nol2","bench_walk","bench_walk_nol2"):
    try:
      for l in open("gpurun_out/r02i_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["config"]["ms_breakdown"], d["config"].get("parity_checked"), d["config"]["setup"].get("prefilter",{}).get("variant"))
    except Exception as e: print(f, "ERR", e)
PY
