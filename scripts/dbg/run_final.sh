python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02final_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02final_gputests.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err; echo "bench rc=$?"; python - <<PY
import json
for l in open("gpurun_out/r02final_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench", round(d["value"],1), round(d["ms_per_step"],3), d["roofline"]["frac"], d["config"]["ms_breakdown"], d["e2e"]["value"], d["config"]["parity_checked"])
PY
