timeout 600 python bench.py --no-e2e --no-cpu --no-parity > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r02x_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("chain64", round(d["value"],1), round(d["ms_per_step"],3), d["config"]["ms_breakdown"])
PY
timeout 1200 python scripts/cli_feeder.py 100000 600 16 /dev/shm/yrb_cli100k > gpurun_out/r02x_cli_feeder_100k.jsonl 2> gpurun_out/r02x_cli_feeder_100k.err; echo "cli rc=$?"; cut -c1-230 gpurun_out/r02x_cli_feeder_100k.jsonl; tail -3 gpurun_out/r02x_cli_feeder_100k.err | cut -c1-300
