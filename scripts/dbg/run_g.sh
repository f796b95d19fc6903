for m in 0 6; do
YRB_SORT_CANDS=$m timeout 600 python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02g_cfg5_m$m.json 2> gpurun_out/r02g_cfg5_m$m.err; echo "rc=$?"
done
python - <<PY
import json
for f in ("cfg5_m0","cfg5_m6"):
    for l in open("gpurun_out/r02g_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3), d["config"]["ms_breakdown"])
PY
