#!/bin/bash
# Round-2 evidence on one B200: the -m gpu suite, the default bench line, the ncu launch list of the same
# command and one `ncu --set full` capture of the scan kernel over the 16 GiB block.
# usage (under gpurun): bash scripts/r02_evidence.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${TAG}_gputests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ac_filter -s 3 -c 1 -f -o gpurun_out/${TAG}_filter_16g \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -c 600 gpurun_out/${TAG}_bench_n1.json
