timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -m gpu -x -q -k "synthetic or config or known_answer or sharded or engine or gather or device_scan" > gpurun_out/r02q_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02q_gputests.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "rc=$?"
YR_B200_GATHER_MATCH_DATA=0 timeout 900 python bench.py --no-cpu --no-parity > gpurun_out/r02q_bench_nogather.json 2> gpurun_out/r02q_bench_nogather.err; echo "rc=$?"
python - <<PY
import json
for f in ("bench","bench_nogather"):
    for l in open("gpurun_out/r02q_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), d["e2e"], d["config"].get("parity_checked"))
PY
