#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of
the pipeline on a few MiB, both scan kernels, shards, multi-block with chains."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import yrh
from yara_b200 import Engine, DeviceBuffer, shard, synth, api

rs = synth.ruleset(4, 1500)
rb = yrh.b200().compile(rs.source)
image = rb.flatten()
data = synth.corpus(rs, (1 << 21) + 333, 2048)
eng = Engine(image, 0)
want = None
for kernel in (api.KERNEL_FILTER, api.KERNEL_WALK):
    eng.set_option(api.OPT_SCAN_KERNEL, kernel)
    r = eng.scan_host(data)
    assert r.rc == 0 and r.error == 0
    t = sorted(zip(r.matches["string_idx"].tolist(), r.matches["offset"].tolist(), r.matches["match_length"].tolist()))
    assert want is None or t == want
    want = t
sh = shard.scan_sharded(eng, data, 3)
assert sorted(zip(sh.matches["string_idx"].tolist(), sh.matches["offset"].tolist(), sh.matches["match_length"].tolist())) == want
b = rb.scan_blocks(data[: 1 << 16], 1024, 256)
assert b.rc == 0
print("sanitize smoke ok:", len(want), "matches")
