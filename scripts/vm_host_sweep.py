#!/usr/bin/env python3
"""CPU-only sweep: the product's verifier device code replayed on the CPU (tests/harness/vm_host.cpp)
vs the oracle's match-callback trace, over tests/fuzzgen.py seeds.  See tests/test_vm_host_cpu.py.

  python scripts/vm_host_sweep.py FIRST LAST [skip,seeds] [scan]
      skip: seeds on which the reference/oracle hangs
      scan: go on through the resolve stage and compare final + unconfirmed matches (normal and fast mode)
            instead of the callback trace"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzzgen, yrh
import test_vm_host_cpu as T

first, last = int(sys.argv[1]), int(sys.argv[2])
skip = set(int(x) for x in sys.argv[3].split(",") if x) if len(sys.argv) > 3 else set()
scan_mode = len(sys.argv) > 4 and sys.argv[4] == "scan"
R = yrh.ref()
bad, done, recs, t0 = [], 0, 0, time.time()
tiers = [0, 0, 0, 0]
for seed in range(first, last):
    if seed in skip:
        continue
    src, insts = fuzzgen.gen_rules(seed, 50)
    good = []
    for s in src:
        try:
            R.compile(s); good.append(s)
        except ValueError:
            pass
    image = R.compile("".join(good)).flatten()
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    try:
        if scan_mode:
            recs += T.check_scan(image, data, 0)
            T.check_scan(image, data, 1)
        else:
            n, st = T.check(image, data)
            recs += n
            for k, f in enumerate(("retry_fast", "retry_re", "retry_fast3", "retry_fast4")):
                tiers[k] += getattr(st, f)
    except BaseException as e:
        if type(e).__name__ == "Skipped":
            done += 1
            continue
        if not isinstance(e, AssertionError):
            raise
        bad.append(seed)
        print("MISMATCH seed", seed, str(e)[:200], flush=True)
    done += 1
if scan_mode:
    print("verify + resolve code on CPU vs oracle, seeds %d..%d: %d done, %d mismatches %s; %d final matches compared; %.0f s"
          % (first, last, done, len(bad), bad, recs, time.time() - t0))
else:
    print("verifier code on CPU vs oracle, seeds %d..%d: %d done, %d mismatches %s; %d callbacks compared; candidates sent to tier 2 fast/re, 3, 4: %s; %.0f s"
          % (first, last, done, len(bad), bad, recs, tiers, time.time() - t0))
