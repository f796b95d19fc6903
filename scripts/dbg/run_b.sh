timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02b_gputests.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r02b_bench_side.json 2> gpurun_out/r02b_bench_side.err; echo "rc=$?"
YRB_SIDE_STREAMS=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02b_bench_noside.json 2> gpurun_out/r02b_bench_noside.err; echo "rc=$?"
python - <<PY
import json
for f in ("side","noside"):
    for l in open("gpurun_out/r02b_bench_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["ms_breakdown"], d["config"].get("parity_checked"))
PY
bash scripts/r02_other_configs.sh gpurun_out/r02b_other_configs.jsonl
