#!/bin/bash
# Round-2 evidence, second part: the other configs, the whole-block parity run, upload timing, the fiber VM profile.
TAG=${1:-r02}
mkdir -p gpurun_out
bash scripts/r02_other_configs.sh gpurun_out/${TAG}_other_configs.jsonl
timeout 1500 python bench.py --no-e2e --no-cpu --parity-full > gpurun_out/${TAG}_full_parity_16gib.json 2> gpurun_out/${TAG}_full_parity_16gib.err; echo "parity-full rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_full_parity_16gib.json"):
    if l.startswith("{"):
        d=json.loads(l); p=d["config"]["parity"]
        print("full parity:", {k:p.get(k) for k in ("ok","n_windows","window_bytes","covers_whole_block","reference_matches_all_windows","whole_block_matches","seconds")}, "bad windows:", len([w for w in p["windows"] if not (w["window_equals_reference"] and w["timed_scan_equals_window"])]))
PY
rm -f /tmp/yrb_upload_20000.yarc; rm -rf ~/.cache/yara_b200
python scripts/upload_timing.py 20000 > gpurun_out/${TAG}_upload_timing.jsonl 2>/dev/null
python scripts/upload_timing.py 20000 >> gpurun_out/${TAG}_upload_timing.jsonl 2>/dev/null
cat gpurun_out/${TAG}_upload_timing.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_re_kernel -c 1 -f -o gpurun_out/${TAG}_verify_re_final \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/${TAG}_ncu_re.log 2>&1; echo "ncu rc=$?"
