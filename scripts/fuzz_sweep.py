#!/usr/bin/env python3
"""Differential fuzz sweep on a GPU box: drop-in vs reference library over many seeds
(tests/fuzzgen.py), in four modes -- one block, fast mode, 1024-byte blocks, and as
items of a file batch (yr_b200_batch_scan + per-buffer scans).  Prints every seed that
differs; exit code 1 if any did.

  python scripts/fuzz_sweep.py FIRST LAST [time-limit-seconds] [skip,seeds] [cpu]

With `cpu` as the 5th argument the drop-in is the TEST stand-in build (tests/harness/libyara_b200cpu.so: the
same glue objects, the product's verify / resolve device code replayed on the CPU) -- no GPU needed.

Seeds on which the REFERENCE does not terminate (yr_re_exec cycles on zero-width loops, DESIGN.md
section 7) must be skipped: pre-screen with the reference alone and pass them as the 4th argument."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fuzzgen, yrh

first, last = int(sys.argv[1]), int(sys.argv[2])
limit = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
skip = set(int(x) for x in sys.argv[4].split(",") if x) if len(sys.argv) > 4 else set()
CPU = len(sys.argv) > 5 and sys.argv[5] == "cpu"
R, B = yrh.ref(), (yrh.b200cpu() if CPU else yrh.b200())
bad, done, t0, slow = [], 0, time.time(), []
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "fuzz_sweep.log"), "a")

def note(*a):
    print(*a, flush=True)
    print(*a, file=log, flush=True)

def same(a, b):
    if a.rc != b.rc:
        return False
    return a.rc != 0 or (a.matches.tobytes() == b.matches.tobytes() and a.rule_bits.tobytes() == b.rule_bits.tobytes())

for seed in range(first, last):
    if time.time() - t0 > limit:
        break
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
    rr, rb = R.compile(text), B.compile(text)
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    t1 = time.time()
    why = []
    note("seed", seed, "start")
    for flags in (0, 1):
        if not same(rr.scan(data, flags=flags, cap=1 << 20), rb.scan(data, flags=flags, cap=1 << 20)):
            why.append("single flags=%d" % flags)
    note("  single scans %.2f s" % (time.time() - t1))
    if time.time() - t1 > 4:      # pathological for the chain replay: do not multiply it by 12
        slow.append((seed, round(time.time() - t1, 1)))
        if why:
            bad.append((seed, why)); note("MISMATCH seed", seed, why)
        done += 1
        continue
    small = data[: 1 << 14]
    if not same(rr.scan_blocks(small, 1024, 256, cap=1 << 20), rb.scan_blocks(small, 1024, 256, cap=1 << 20)):
        why.append("blocks")
    # the same data cut into ragged items, scanned as one batch
    cuts = sorted(set([0, len(data)] + [(seed * 7919 + k * 10007) % len(data) for k in range(9)]))
    items = [np.ascontiguousarray(data[a:b]) for a, b in zip(cuts, cuts[1:])] + [np.ascontiguousarray(data)]
    want = [rr.scan(it, cap=1 << 20) for it in items]
    with rb.batch(items) as bt:
        if bt.rc != 0:
            why.append("batch rc=%d" % bt.rc)
        else:
            for k, it in enumerate(items):
                if not same(want[k], rb.scan(it, cap=1 << 20)):
                    why.append("batch item %d" % k)
                    break
    dt = time.time() - t1
    if dt > 5:
        slow.append((seed, round(dt, 1)))
    if why:
        bad.append((seed, why))
        note("MISMATCH seed", seed, why)
    note("  all modes %.2f s" % dt)
    done += 1
note("seeds %d..%d: %d done, %d mismatches %s; slow: %s; %.0f s" % (first, first + done, done, len(bad), bad, slow, time.time() - t0))
sys.exit(1 if bad else 0)
