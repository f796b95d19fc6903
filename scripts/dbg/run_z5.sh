timeout 1500 python scripts/cli_feeder.py 100000 600 16 /dev/shm/yrb_cli100k > gpurun_out/r02z5_cli_feeder_100k.jsonl 2> gpurun_out/r02z5_cli_feeder_100k.err; echo "cli rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r02z5_cli_feeder_100k.jsonl"):
    d=json.loads(l)
    if "arm" in d: print(d["arm"], "| total", d["seconds"], "s | one file", d["one_file_process_seconds"], "s | beyond startup", d["beyond_startup"]["seconds"], "s =", d["beyond_startup"]["files_per_s"], "files/s", d["beyond_startup"]["GB_per_s"], "GB/s")
    else: print(d)
PY
