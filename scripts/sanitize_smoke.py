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

# ---- round 2: the other filter-kernel variants, the fiber VM's local tier + sorted candidates, packed match data
SHORT = 'rule sa { strings: $mz = "MZ" $pk = { 50 4B 03 } $q = "q" fullword condition: any of them }\n'
for name, src, every in (("sparse (variant 3)", synth.ruleset(2, 300).source, 4096),
                         ("short atoms (variant 6)", SHORT + synth.ruleset(3, 400).source, 4096),
                         ("regex (variant 5 or 3, fiber VM tier 0, sorted candidates)", synth.ruleset(5, 1500).source, 512)):
    rs2 = synth.ruleset(5, 1500) if "regex" in name else (synth.ruleset(2, 300) if "sparse" in name else synth.ruleset(3, 400))
    rb2 = yrh.b200().compile(src)
    e2 = Engine(rb2.flatten(), 0)
    d2 = synth.corpus(rs2, (1 << 21) + 777, every)
    got = None
    for kernel in (api.KERNEL_FILTER, api.KERNEL_WALK):
        e2.set_option(api.OPT_SCAN_KERNEL, kernel)
        r = e2.scan_host(d2)
        assert r.rc == 0 and r.error == 0, (name, r.rc, r.error)
        t = sorted(zip(r.matches["string_idx"].tolist(), r.matches["offset"].tolist(), r.matches["match_length"].tolist()))
        assert got is None or t == got, name
        got = t
    print("sanitize smoke ok:", name, e2.filter_info()["variant"], len(got), "matches", r.stats["n_candidates"], "candidates")
many = np.frombuffer(b"abcd ef123gh " * 9000, dtype=np.uint8)
rm = yrh.b200().compile('rule t { strings: $a = "abcd" $b = /ef[0-9]{1,6}gh/ condition: any of them }')
assert len(rm.scan(many, cap=1 << 18).matches) == 18000          # >= 8192 matches: yrb_last_match_data
assert rm.scan_blocks(many[: 40000], 1024, 256, cap=1 << 18).rc == 0   # 52 small blocks: one engine batch
print("sanitize smoke ok: packed match data, block batch")
