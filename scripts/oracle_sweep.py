#!/usr/bin/env python3
"""CPU-only differential sweep: the oracle restatement (oracle/yr_oracle.c) vs the reference library
built from /root/reference (oracle/_ref), over many tests/fuzzgen.py seeds, normal and fast mode.
Seeds on which the reference does not terminate must be skipped.

  python scripts/oracle_sweep.py FIRST LAST [skip,seeds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fuzzgen, yrh

first, last = int(sys.argv[1]), int(sys.argv[2])
skip = set(int(x) for x in sys.argv[3].split(",") if x) if len(sys.argv) > 3 else set()
R = yrh.ref()
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
    r = R.compile("".join(good))
    img = r.flatten()
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    for flags in (0, 1):
        ref = r.scan(data, flags=flags, cap=1 << 20)
        o = yrh.oracle_scan(img, data, flags=flags)
        ok = o.error == ref.rc
        if ok and ref.rc == 0:
            got = yrh.matches_from_flat(o.matches, data)
            if len(got) == len(ref.matches):
                got["is_private"] = ref.matches["is_private"]
            ok = got.tobytes() == ref.matches.tobytes()
        if not ok:
            bad.append((seed, flags, ref.rc, o.error, len(ref.matches), len(o.matches)))
            print("MISMATCH", bad[-1], flush=True)
    done += 1
print("oracle vs reference, seeds %d..%d: %d done, %d mismatches %s; %.0f s" % (first, last, done, len(bad), bad, time.time() - t0))
