python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02c_gputests.log
YRB_TRACE=1 timeout 600 python scripts/fuzz_sweep.py 0 120 400 > gpurun_out/r02c_fuzz.log 2> gpurun_out/r02c_fuzz.err; echo "fuzz rc=$?"; tail -2 gpurun_out/r02c_fuzz.log; grep -c "plain records deduplicated" gpurun_out/r02c_fuzz.err
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "rc=$?"
timeout 600 python bench.py --no-e2e --no-cpu --size-gib 1 --parity-full > gpurun_out/r02c_bench_pf1.json 2> gpurun_out/r02c_bench_pf1.err; echo "rc=$?"
python - <<PY
import json
for f in ("r02c_bench","r02c_bench_pf1"):
    for l in open("gpurun_out/%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l); p=d["config"]["parity"] or {}
            print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["ms_breakdown"], d["config"].get("parity_checked"), {k:p.get(k) for k in ("n_windows","covers_whole_block","reference_matches_all_windows","seconds","ok")})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02c_ncu_launches_cfg5.csv \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02c_ncu_cfg5.log 2>&1; echo "launch list rc=$?"
