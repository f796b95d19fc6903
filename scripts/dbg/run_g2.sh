YRB_SORT_CANDS=6 timeout 600 python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02g_cfg5_nofast.json 2> gpurun_out/r02g_cfg5_nofast.err; echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r02g_cfg5_nofast.json"):
    if l.startswith("{"):
        d=json.loads(l)
        print("nofast", round(d["value"],1), round(d["ms_per_step"],3), d["config"]["ms_breakdown"])
PY
