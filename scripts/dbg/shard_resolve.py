"""single GPU: cost of the pieces of the sharded flow (defer + publish + resolve_raw_device) vs a plain scan"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from yara_b200 import Engine, LibYara, synth, api
sys.argv = sys.argv[:1]
import bench
ly = LibYara(); rs = synth.ruleset(3, 20000); rules = ly.compile(rs.source); image = ly.flatten(rules)
eng = Engine(image, 0)
S = 8 << 30
dbuf, n = bench.make_shard(api, synth, rs, 0, S, 0, S, 65536, 8, True)
for _ in range(3): r = eng.scan_device(dbuf, copy=False)
t0 = time.perf_counter(); r = eng.scan_device(dbuf, copy=False); t1 = time.perf_counter()
print("plain scan: wall %.3f ms, stats %s" % ((t1 - t0) * 1e3, {k: round(v, 3) for k, v in r.stats.items() if k.startswith("ms_")}))
eng.shard_export_create(1 << 22)
ptr = eng.shard_export_ptr()
for it in range(4):
    t0 = time.perf_counter()
    r = eng.scan_device(dbuf, copy=False, data_offset=0, block_size=S, own=(0, S), defer=True)
    t1 = time.perf_counter()
    n, err = eng.shard_export_publish()
    t2 = time.perf_counter()
    res = eng.resolve_raw_device([ptr], [n], copy=False)
    t3 = time.perf_counter()
    print("defer scan %.3f ms (stats %s), publish %.3f ms (%d records), resolve_raw_device %.3f ms (event %.3f, %d matches)" % (
        (t1 - t0) * 1e3, {k: round(v, 3) for k, v in r.stats.items() if k.startswith("ms_")}, (t2 - t1) * 1e3, n, (t3 - t2) * 1e3,
        res.stats["ms_resolve"], res.n_matches))
