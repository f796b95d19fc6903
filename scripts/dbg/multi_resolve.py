"""2 GPUs, one process: yrb_multi_scan_device timing (gather by cudaMemcpyPeerAsync + global resolve)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from yara_b200 import Engine, LibYara, synth, api
from yara_b200.api import MultiEngine
sys.argv = sys.argv[:1]
import bench
ly = LibYara(); rs = synth.ruleset(3, 20000); rules = ly.compile(rs.source); image = ly.flatten(rules)
nd = Engine.device_count()
S = 16 << 30
m = MultiEngine(image, list(range(nd)))
plan = m.plan(S)
bufs = [bench.make_shard(api, synth, rs, g, S, plan[g][0], plan[g][1], 65536, 8, True)[0] for g in range(nd)]
for it in range(6):
    t0 = time.perf_counter()
    r = m.scan_device(bufs, plan, S, copy=False)
    t1 = time.perf_counter()
    print("multi scan_device: wall %.3f ms rc %d matches %d stats %s" % ((t1 - t0) * 1e3, r.rc, r.n_matches, {k: round(v, 3) for k, v in r.stats.items() if k.startswith("ms_")}))
