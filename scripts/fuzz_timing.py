#!/usr/bin/env python3
"""Timing of one fuzz seed through the drop-in (YRB_TIMING=1 prints the engine's phase times)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import yrh, fuzzgen
seed = int(sys.argv[1])
R, B = yrh.ref(), yrh.b200()
text, insts, n = fuzzgen.compile_filtered(R, seed, 60)
rr, rb = R.compile(text), B.compile(text)
data = fuzzgen.gen_data(seed, insts, (1 << 17) + 13 * seed)
t = time.time(); a = rr.scan(data, cap=1 << 20); tr = time.time() - t
o = yrh.oracle_scan(rr.flatten(), data)
print("seed", seed, "ref %.3f s" % tr, "matches", len(a.matches), "candidates", o.n_candidates, "maxpos", o.max_positions, "maxfib", o.max_fibers)
t = time.time(); b = rb.scan(data, cap=1 << 20); tb = time.time() - t
print("gpu %.3f s" % tb, "same", a.matches.tobytes() == b.matches.tobytes())
