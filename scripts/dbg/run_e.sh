timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r02e_gputests.log
timeout 600 python scripts/fuzz_sweep.py 0 200 300 > gpurun_out/r02e_fuzz.log 2> gpurun_out/r02e_fuzz.err; echo "fuzz rc=$?"; tail -1 gpurun_out/r02e_fuzz.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "rc=$?"
bash scripts/r02_other_configs.sh gpurun_out/r02e_other_configs.jsonl
YRB_FILTER_NO_DENSE=1 timeout 600 python bench.py --config 4 --strings 100000 --every 65536 --size-gib 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02e_bench_cfg4_nodense.json 2> gpurun_out/r02e_bench_cfg4_nodense.err; echo "rc=$?"
python - <<PY
import json
for f in ("r02e_bench","r02e_bench_cfg4_nodense"):
    for l in open("gpurun_out/%s.json"%f):
        if l.startswith("{"):
            d=json.loads(l); p=d["config"]["parity"] or {}
            print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["ms_breakdown"], d["config"].get("parity_checked"), d["config"]["setup"].get("prefilter"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_re_kernel -c 1 -f -o gpurun_out/r02e_verify_re \
  python bench.py --config 5 --strings 5000 --every 16384 --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02e_ncu_re.log 2>&1; echo "ncu rc=$?"
