timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02d_gputests.log
timeout 600 python scripts/fuzz_sweep.py 0 150 300 > gpurun_out/r02d_fuzz.log 2> gpurun_out/r02d_fuzz.err; echo "fuzz rc=$?"; tail -1 gpurun_out/r02d_fuzz.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "rc=$?"
timeout 600 python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --no-e2e --no-cpu > gpurun_out/r02d_bench_cfg5.json 2> gpurun_out/r02d_bench_cfg5.err; echo "rc=$?"
YRB_RE_TIER0=0 timeout 600 python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02d_bench_cfg5_not0.json 2> gpurun_out/r02d_bench_cfg5_not0.err; echo "rc=$?"
python - <<PY
import json
for f in ("r02d_bench","r02d_bench_cfg5","r02d_bench_cfg5_not0"):
    for l in open("gpurun_out/%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l); p=d["config"]["parity"] or {}
            print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["ms_breakdown"], d["config"].get("parity_checked"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_re_kernel -c 2 -f -o gpurun_out/r02d_verify_re \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02d_ncu_re.log 2>&1; echo "ncu rc=$?"
YRB_RE_TIER0=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_re_kernel -c 1 -f -o gpurun_out/r02d_verify_re_not0 \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02d_ncu_re_not0.log 2>&1; echo "ncu rc=$?"
