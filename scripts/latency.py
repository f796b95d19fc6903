#!/usr/bin/env python3
"""Per-scan latency of small blocks through the drop-in (file-batch shape, BASELINE configs[3])."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import yrh
from yara_b200 import synth

rs = synth.ruleset(4, int(sys.argv[1]) if len(sys.argv) > 1 else 20000)
rb = yrh.b200().compile(rs.source)
rr = yrh.ref().compile(rs.source)
for size in (65536, 1 << 20, 8 << 20):
    data = synth.corpus(rs, size, 65536)
    rb.scan(data)
    t = time.time(); n = 50
    for _ in range(n): rb.scan(data)
    tb = (time.time() - t) / n
    t = time.time(); n2 = 5
    for _ in range(n2): rr.scan(data)
    tr = (time.time() - t) / n2
    print("block %8d B: drop-in %.3f ms/scan (%.2f GB/s), reference 1 thread %.3f ms/scan (%.3f GB/s)" % (size, tb * 1e3, size / tb / 1e9, tr * 1e3, size / tr / 1e9))
