#!/usr/bin/env python3
"""CPU-only sweep: the whole drop-in (reference host code + hooks + host/yr_b200_glue.c) linked against the
TEST stand-in engine (tests/harness/libyara_b200cpu.so) vs the reference library, over tests/fuzzgen.py
seeds -- the per-seed body of tests/test_glue_cpu.py::test_fuzz_rulesets_through_the_glue: whole-block scans
in normal and fast mode, 1024-byte blocks (chains across blocks: seeds; small blocks deferred and batched),
and the same bytes cut into a parked file batch.

  python scripts/glue_sweep.py FIRST LAST [skip,seeds]

A seed on which the reference never returns (deviation 0) hangs this script too: run it
under `timeout` over short ranges and list such seeds in `skip`."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_glue_cpu as T

first, last = int(sys.argv[1]), int(sys.argv[2])
skip = set(int(x) for x in sys.argv[3].split(",") if x) if len(sys.argv) > 3 else set()
bad, done, t0 = [], 0, time.time()
for seed in range(first, last):
    if seed in skip:
        continue
    t1 = time.time()
    try:
        T.run_fuzz([seed])
    except AssertionError as e:
        bad.append(seed)
        print("MISMATCH seed", seed, str(e)[:300], flush=True)
    done += 1
    if time.time() - t1 > 30:
        print("slow seed %d: %.0f s" % (seed, time.time() - t1), flush=True)
print("drop-in on the CPU stand-in vs reference library, seeds %d..%d: %d done, %d mismatches %s; %.0f s"
      % (first, last, done, len(bad), bad, time.time() - t0))
