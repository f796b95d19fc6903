#!/bin/bash
# HBM-resident throughput of the non-headline ruleset shapes (parity-test cases, not bench lines)
run() { python bench.py --config $1 --strings $2 --size-gib $3 --every $4 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']; print('config $1 strings $2 ${5}:', 'value %.0f GB/s'%d['value'], 'kernel', c['scan_kernel'], 'scan %.0f GB/s'%d['roofline']['achieved'], c['ms_breakdown'], 'matches', c['matches_rank0'], 'cands', c['candidates_rank0'])"; }
run 2 1000 1 65536
run 5 5000 1 16384
run 4 100000 1 65536
YRB_FILTER_MAX_FILL=0.7 run 4 100000 1 65536 "(filter forced)"
YRB_SCAN_KERNEL=1 run 5 5000 1 16384 "(walk)"
