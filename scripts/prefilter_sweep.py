#!/usr/bin/env python3
"""CPU-only differential sweep of the condition-VM pre-filter: reference library vs
oracle/_ref/libyara_refpf.so (reference + pre-filter hooks, filter ON) over tests/fuzzgen.py seeds;
every ruleset is scanned repeatedly so that learnt skips are exercised.

  python scripts/prefilter_sweep.py FIRST LAST [skip,seeds]   (skip: seeds on which the reference hangs)"""
import os, sys, time
os.environ["YR_B200_RULE_PREFILTER"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fuzzgen, yrh
from yara_b200 import synth

first, last = int(sys.argv[1]), int(sys.argv[2])
skip = set(int(x) for x in sys.argv[3].split(",") if x) if len(sys.argv) > 3 else set()
R, P = yrh.ref(), yrh.refpf()
bad, done, t0 = [], 0, time.time()
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
    text = "".join(good)
    rr, rp = R.compile(text), P.compile(text)
    data = np.frombuffer(bytes(fuzzgen.gen_data(seed, insts, 1 << 15)), dtype=np.uint8)
    noise = synth.rand_bytes(seed, 0, 2048).copy()
    for k, buf in enumerate((noise, data, noise, np.ascontiguousarray(data[: len(data) // 3]), data, np.zeros(0, np.uint8))):
        for flags in (0, 1):
            a, b = rr.scan(buf, flags=flags, cap=1 << 20), rp.scan(buf, flags=flags, cap=1 << 20)
            if a.rc != b.rc or a.rule_bits.tobytes() != b.rule_bits.tobytes() or a.matches.tobytes() != b.matches.tobytes():
                bad.append((seed, k, flags))
                print("MISMATCH", bad[-1], flush=True)
    done += 1
e, s = yrh.prefilter_stats(P)
print("pre-filter vs reference, seeds %d..%d: %d done, %d mismatches %s; rules evaluated %d, skipped %d; %.0f s"
      % (first, last, done, len(bad), bad, e, s, time.time() - t0))
