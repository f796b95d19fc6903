#!/usr/bin/env python3
"""BASELINE configs[3] shape: a batch of blobs (log-uniform 64 KiB - 8 MiB), one
yr_rules_scan_mem per blob, N scanner threads over one shared YR_RULES -- the
structure of `yara -p N` (cli/yara.c:1561-1610).  Compares the GPU drop-in with
the reference library on the same host threads.  Blobs live in host memory."""
import sys, time, math, random
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import yrh
from yara_b200 import synth

nstrings = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 600
nthreads = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["1", "8", "16"])]
no_batch = "--no-batch" in sys.argv

rs = synth.ruleset(4, nstrings)
rng = random.Random(3)
sizes = [int(math.exp(rng.uniform(math.log(65536), math.log(8 << 20)))) for _ in range(nfiles)]
blobs = [synth.corpus(rs, s, 65536, seed=1000 + k, plant_seed=k, tail_plant=False) for k, s in enumerate(sizes)]
total = sum(sizes)
print("ruleset: %d declared strings; %d blobs, %.2f GiB, mean %.2f MiB" % (nstrings, nfiles, total / 2**30, total / nfiles / 2**20))

def run(rules, n):
    def one(b):
        r = rules.scan(b, cap=4096)
        assert r.rc == 0
        return len(r.matches)
    t = time.time()
    with ThreadPoolExecutor(n) as ex:
        m = sum(ex.map(one, blobs))
    return time.time() - t, m

rb = yrh.b200().compile(rs.source)
rr = yrh.ref().compile(rs.source)
run(rb, max(nthreads))  # warm-up: one pooled device scanner per thread (creating one costs ~50 ms)

# ---- the same blobs as file batches (yrb_scan_batch / yr_b200_batch_scan)
from yara_b200 import Engine, PinnedBuffer
image = rb.flatten()
eng = Engine(image, 0)
pin = PinnedBuffer(sum((s + 15) & ~15 for s in sizes))
views, o = [], 0
for b in blobs:
    pin.array[o : o + len(b)] = b
    views.append(pin.array[o : o + len(b)])
    o += (len(b) + 15) & ~15
for label, its, one in (() if no_batch else (("pageable, packed", blobs, False), ("pinned, packed", views, False), ("pinned, one allocation", views, True))):
    for per in (len(its), 64):
        eng.scan_batch(its[:per], one_allocation=one)
        t = time.time(); nm = nr = nl = 0
        for k in range(0, len(its), per):
            r = eng.scan_batch(its[k : k + per], one_allocation=one)
            assert r.rc == 0
            nm += len(r.all); nr += r.n_rescan; nl += r.stats["kernel_launches"]
        t = time.time() - t
        print("engine batch (%s, %d items/batch): %.3f s  %.2f GB/s  %.0f files/s  (%d matches, %d rescans, %d launches)"
              % (label, per, t, total / t / 1e9, nfiles / t, nm, nr, nl))
eng.close()
for n in ([] if no_batch else nthreads):
    t = time.time()
    with rb.batch(views, one_allocation=True) as bt:
        assert bt.rc == 0
        tb0 = time.time() - t
        def one(b):
            r = rb.scan(b, cap=4096)
            assert r.rc == 0
            return len(r.matches)
        with ThreadPoolExecutor(n) as ex:
            m = sum(ex.map(one, views))
    t = time.time() - t
    print("drop-in batch + %2d scan threads: %.2f s (batch %.2f s)  %.2f GB/s  %.0f files/s  (%d matches, %d rescans)"
          % (n, t, tb0, total / t / 1e9, nfiles / t, m, bt.n_rescans))
for n in nthreads:
    tb, mb = run(rb, n)
    print("drop-in   %2d threads: %.2f s  %.2f GB/s  %.0f files/s  (%d matches)" % (n, tb, total / tb / 1e9, nfiles / tb, mb))
sub = blobs[: max(8, nfiles // 20)]
tsub = sum(len(b) for b in sub)
blobs_all, blobs = blobs, sub
for n in (nthreads[-1],):
    tr, mr = run(rr, n)
    print("reference %2d threads: %.2f s  %.3f GB/s  %.0f files/s on a %d-blob subset (%d matches)" % (n, tr, tsub / tr / 1e9, len(sub) / tr, len(sub), mr))
