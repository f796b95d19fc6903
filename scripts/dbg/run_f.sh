timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or config5 or synthetic" > gpurun_out/r02f_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02f_gputests.log
for m in 0 4 6; do
YRB_SORT_CANDS=$m timeout 600 python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02f_cfg5_m$m.json 2> gpurun_out/r02f_cfg5_m$m.err; echo "rc=$?"
YRB_SORT_CANDS=$m timeout 600 python bench.py --size-gib 4 --no-e2e --no-cpu --no-parity > gpurun_out/r02f_cfg3_m$m.json 2> gpurun_out/r02f_cfg3_m$m.err; echo "rc=$?"
done
python - <<PY
import json
for f in ("cfg5_m0","cfg5_m4","cfg5_m6","cfg3_m0","cfg3_m4","cfg3_m6"):
    for l in open("gpurun_out/r02f_%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l)
            print(f, round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3), d["config"]["ms_breakdown"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_re_kernel -c 1 -f -o gpurun_out/r02f_verify_re \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02f_ncu_re.log 2>&1; echo "ncu rc=$?"
