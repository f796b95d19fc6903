#!/bin/bash
# Driver-style bench lines for the other BASELINE configs and for low-entropy corpora, 1 GiB, one B200.
# usage (under gpurun): bash scripts/r02_other_configs.sh [out.jsonl]
OUT=${1:-gpurun_out/r02_other_configs.jsonl}
: > $OUT
run() { echo "# $*" >&2; timeout 900 python bench.py --size-gib 1 --steps 5 --warmup 3 --no-e2e --no-cpu "$@" >> $OUT 2>> gpurun_out/other_configs.err; }
run --config 2 --strings 1000
run --config 4 --strings 100000 --every 65536
run --config 5 --strings 5000 --every 16384
run --config 3 --strings 20000 --corpus zero
run --config 3 --strings 20000 --corpus text
python - <<PY
import json
for l in open("$OUT"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l); c=d["config"]
    print("%-70s value %8.1f GB/s  kernel %8.1f (%.3f)  %s  breakdown %s parity %s" % (c["workload"][:70], d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], c["scan_kernel"]+"/"+c["corpus"], c["ms_breakdown"], c["parity_checked"]))
PY
