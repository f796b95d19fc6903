#!/usr/bin/env python3
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import yrh
from yara_b200 import synth
rs = synth.ruleset(4, 20000)
rb = yrh.b200().compile(rs.source)
data = synth.corpus(rs, 65536, 65536)
for _ in range(3): rb.scan(data)
import os
os.environ["YRB_TIMING"] = "1"
t = time.time()
for _ in range(5): rb.scan(data)
print("wall per scan %.3f ms" % ((time.time() - t) / 5 * 1e3))
