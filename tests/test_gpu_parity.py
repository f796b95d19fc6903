"""GPU parity tests (-m gpu): the drop-in libyara whose block scan runs on the
B200 (yara_b200/libyara_b200.so -> yr_b200_scan_mem_block -> C ABI ->
libyrb200.so) against the reference libyara built from /root/reference
(oracle/_ref/libyara_ref.so) through the SAME public API (yr_rules_scan_mem /
yr_scanner_scan_mem_blocks), bit-exact on every YR_MATCH field the harness
dumps: (string_idx, base, offset, match_length, data_length, hash(data),
xor_key, is_private), the rule verdicts, the return code and the number of
TOO_MANY_MATCHES callbacks."""
import os

import numpy as np
import pytest

import yrh
from cases import CASES
from yara_b200 import synth

pytestmark = pytest.mark.gpu


def both(rule_text):
    return yrh.ref().compile(rule_text), yrh.b200().compile(rule_text)


def assert_same(a, b):
    assert a.rc == b.rc
    assert len(a.matches) == len(b.matches), (len(a.matches), len(b.matches))
    if a.matches.tobytes() != b.matches.tobytes():
        for x, y in zip(a.matches, b.matches):
            assert x == y, (x, y)
    assert a.rule_bits.tobytes() == b.rule_bits.tobytes()
    assert a.too_many_calls == b.too_many_calls


def test_engine_loaded():
    import ctypes
    lib = ctypes.CDLL(os.path.join(yrh.ROOT, "yara_b200", "libyrb200.so"))
    assert lib.yrb_device_count() >= 1


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases(idx):
    rule, data = CASES[idx]
    rr, rb = both(rule)
    assert_same(rr.scan(data), rb.scan(data))


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases_fast_mode(idx):
    rule, data = CASES[idx]
    rr, rb = both(rule)
    assert_same(rr.scan(data, flags=1), rb.scan(data, flags=1))


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases_blocks(idx):
    """tests/test-rules.c pass 3: 1024-byte blocks with 256 bytes of overlap."""
    rule, data = CASES[idx]
    rr, rb = both(rule)
    assert_same(rr.scan_blocks(data, 1024, 256), rb.scan_blocks(data, 1024, 256))


def test_empty_and_tiny_blocks():
    rr, rb = both('rule t { strings: $a = "abc" $b = /x*/ condition: any of them }')
    for data in (b"", b"a", b"ab", b"abc", b"xabc"):
        assert_same(rr.scan(data), rb.scan(data))


@pytest.mark.parametrize("cfg,n,size,every", [
    (2, 1000, 1 << 24, 4096),
    (3, 20000, 1 << 24, 2048),
    (4, 20000, 1 << 23, 2048),
    (5, 5000, 1 << 23, 1024),
])
def test_synthetic_configs(cfg, n, size, every):
    rs = synth.ruleset(cfg, n)
    rr, rb = both(rs.source)
    data = synth.corpus(rs, size, every)
    a = rr.scan(data, cap=1 << 20)
    b = rb.scan(data, cap=1 << 20)
    assert len(a.matches) > size // every // 2
    assert_same(a, b)


def test_synthetic_random_sizes():
    """ragged sizes around the kernels' chunk/vector boundaries"""
    rs = synth.ruleset(3, 800)
    rr, rb = both(rs.source)
    base = synth.corpus(rs, 70000, 256)
    for size in (1, 3, 15, 16, 17, 255, 256, 257, 4095, 4097, 32768 + 5, 65536 + 33, 70000):
        for start in (0, 1, 7):
            data = base[start : start + size].copy()
            assert_same(rr.scan(data), rb.scan(data))


def test_too_many_matches():
    """test-api.c:320-395"""
    rr, rb = both('rule t { strings: $a = "aa" condition: $a }')
    data = np.full(1_200_000, ord("a"), dtype=np.uint8)
    a = rr.scan(data, cap=1_100_000)
    b = rb.scan(data, cap=1_100_000)
    assert a.too_many_calls == 1 and len(a.matches) == 1_000_000
    assert_same(a, b)
    a2 = rr.scan(data, too_many_action=1, cap=16)
    b2 = rb.scan(data, too_many_action=1, cap=16)
    assert a2.rc == 30 and b2.rc == 30


def test_matches_restatement_oracle():
    """third leg: engine == oracle restatement over the flat image"""
    rs = synth.ruleset(3, 3000)
    rb = yrh.b200().compile(rs.source)
    data = synth.corpus(rs, 1 << 22, 2048)
    b = rb.scan(data, cap=1 << 18)
    o = yrh.oracle_scan(rb.flatten(), data)
    got = yrh.matches_from_flat(o.matches, data)
    assert len(got) == len(b.matches)
    for f in ("offset", "string_idx", "match_length", "xor_key", "data_hash"):
        assert (got[f] == b.matches[f]).all(), f


def test_concurrent_scanners_share_rules():
    """`yara -p N` structure (cli/yara.c:1561-1610): N threads, each with its own
    scanner over one shared YR_RULES -> one device rules image, N device scanners."""
    rs = synth.ruleset(3, 2000)
    rr, rb = both(rs.source)
    data = synth.corpus(rs, 1 << 23, 2048)
    secs, rc = rb.scan_mt(data, 8)
    assert rc == 0
    # slices are scanned independently: compare one slice against the reference
    part = data[: len(data) // 8].copy()
    assert_same(rr.scan(part, cap=1 << 18), rb.scan(part, cap=1 << 18))


def test_small_block_latency_smoke():
    """file-batch shaped input (64 KiB blobs): every scan goes through the whole pipeline"""
    rs = synth.ruleset(4, 3000)
    rr, rb = both(rs.source)
    for k in range(20):
        data = synth.corpus(rs, 65536 + 4096 * k, 4096, seed=100 + k, plant_seed=k)
        assert_same(rr.scan(data), rb.scan(data))


def test_more_than_64k_chain_parts_carried_between_blocks():
    """70,000 unconfirmed heads of a chained string at the end of block 1; the tails in block 2 must still
    find them (yrb_scan_params.seeds carries up to 4M nodes; the seed index no longer fits 16 bits)"""
    rule = "rule t { strings: $a = { 11 22 33 44 [300-400] 55 66 77 88 } condition: $a }"
    rr, rb = both(rule)
    head, tail = bytes.fromhex("11223344"), bytes.fromhex("55667788")
    block = 70000 * 4
    b2 = bytearray(b"." * block)
    for off in (350, 100000, 100002, 279000):
        b2[off : off + 4] = tail
    data = np.frombuffer(head * 70000 + bytes(b2), dtype=np.uint8)
    a = rr.scan_blocks(data, block, 0, cap=1 << 20)
    b = rb.scan_blocks(data, block, 0, cap=1 << 20)
    assert a.rc == 0 and len(a.matches) > 20
    assert_same(a, b)


def test_zero_width_loop_that_never_ends_in_the_reference():
    """/^zzyc.((\\b)*?)zbz/ over "zzyc\\0a...": yr_re_exec cycles forever at one input position (the exit
    branch dies on a literal mismatch, the loop branch re-splits into the fiber just freed).  The
    reference only survives such data when an EARLIER candidate fails with ERROR_TOO_MANY_RE_FIBERS;
    the GPU verifies every candidate, cuts the cycle and reports it as the same error (found by
    scripts/fuzz_sweep.py, seed 1002)."""
    rule = r"rule r { strings: $s2 = /^zzyc.((\b)*?)zbz/ condition: $s2 }"
    rr, rb = both(rule)
    # an earlier exploding candidate decides the scan in both implementations
    data = b"xx zzyc zbz ....... zzyc\x00accd0az_zzb1x"
    a, b = rr.scan(data), rb.scan(data)
    assert a.rc == 46 and b.rc == 46
    # the cycling candidate alone: the reference would never return; the GPU path answers
    b = rb.scan(b"zzyc\x00accd0az_zzb1x")
    assert b.rc == 46


def test_repeat_any_recursion_as_deep_as_the_fiber_pool():
    """/a1dx((.?)+yy0)/: _yr_re_fiber_sync recurses once per REPEAT_ANY split (re.c:1626) until the
    1024 fibers are used up -> ERROR_TOO_MANY_RE_FIBERS, not an internal error of the frame stack
    (found by scripts/fuzz_sweep.py, seed 1123)."""
    rr, rb = both("rule r { strings: $s0 = /a1dx((.?)+yy0)/i condition: $s0 }")
    for data in (b"a1dx", b"zza1dxabcyy0 tail", b"nothing here"):
        assert_same(rr.scan(data), rb.scan(data))
    assert rb.scan(b"a1dx").rc == 46


# ---- round 2: the full-size rulesets of BASELINE configs[3] and [4], and the sharded engine

def test_config4_100k_string_ruleset():
    """BASELINE configs[3]: the 100,000-string mixed ruleset (transition table > shared memory, B3 too dense
    for the prefilter -> whatever kernel the engine picks), a single buffer and a 40-item file batch."""
    rs = synth.ruleset(4, 100000)
    rr, rb = both(rs.source)
    assert rr.num_strings >= 100000
    data = synth.corpus(rs, 1 << 23, 2048)
    a = rr.scan(data, cap=1 << 20)
    b = rb.scan(data, cap=1 << 20)
    assert len(a.matches) > 2000
    assert_same(a, b)
    files = [synth.corpus(rs, 65536 + 37 * 1024 * k, 4096, seed=200 + k, plant_seed=k) for k in range(40)]
    want = [rr.scan(f) for f in files]
    with rb.batch(files) as bt:
        assert bt.rc == 0
        for k, f in enumerate(files):
            assert_same(want[k], rb.scan(f))
    for k, f in enumerate(files[:8]):      # and one by one, without the batch
        assert_same(want[k], rb.scan(f))


def test_config5_5000_regexes_64mib():
    """BASELINE configs[4] at its declared size: 5,000 strings (5/6 regex, 1/6 xor/base64 text), 64 MiB with one
    generated instance per 16 KiB so that the fiber VM is really exercised."""
    rs = synth.ruleset(5, 5000)
    rr, rb = both(rs.source)
    data = synth.corpus(rs, 1 << 26, 16384)
    a = rr.scan(data, cap=1 << 20)
    b = rb.scan(data, cap=1 << 20)
    assert len(a.matches) > 3000
    assert_same(a, b)
    assert_same(rr.scan(data[: 1 << 22], flags=1, cap=1 << 20), rb.scan(data[: 1 << 22], flags=1, cap=1 << 20))


def _canon(m):
    """records string-major, ascending offset: the lists of chained strings' heads are appended by one warp
    per chain, in whatever order the warps finish, so only the order INSIDE a string's list is defined"""
    return np.sort(m, order=["string_idx", "offset"]).tobytes()


def _lists_in_order(m):
    """... and inside every string's list the offsets ascend as handed back"""
    last = {}
    for s_, o_ in zip(m["string_idx"].tolist(), m["offset"].tolist()):
        if s_ in last and last[s_] >= o_:
            return False
        last[s_] = o_
    return True


def _corpus_with_boundary_plants(rs, size, every, parts):
    buf = synth.corpus(rs, size, every)
    for off, inst in synth.boundary_plants(rs, size, parts):
        buf[off : off + len(inst)] = np.frombuffer(inst, dtype=np.uint8)
    return buf


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_engine_equals_single_device(shards):
    """yrb_multi_scan_host (one block over several GPUs) == yrb_scan_host on one device ==
    the reference, with instances straddling every shard boundary and chained strings whose parts lie in
    different shards.  Uses as many GPUs as are visible; with fewer than `shards` some get several shards."""
    from yara_b200.api import Engine, MultiEngine
    rs = synth.ruleset(3, 6000)
    rr, rb = both(rs.source)
    size = (48 << 20) + 12288 * shards
    data = _corpus_with_boundary_plants(rs, size, 4096, shards)
    image = rb.flatten()
    ndev = Engine.device_count()
    multi = MultiEngine(image, [g % ndev for g in range(shards)])
    single = Engine(image, 0)
    try:
        per = ((size + shards - 1) // shards + 4095) & ~4095
        assert [p[2] for p in multi.plan(size)] == [min(size, g * per) for g in range(shards)]
        m = multi.scan_host(data)
        s = single.scan_host(data)
        assert m.rc == 0 and s.rc == 0 and m.error == 0
        assert _canon(m.matches) == _canon(s.matches) and len(m.matches) > 10000
        assert _lists_in_order(m.matches) and _lists_in_order(s.matches)
        # matches that straddle a boundary are there, and the reference agrees with all of it
        a = rr.scan(data, cap=1 << 20)
        got = sorted(zip(m.matches["string_idx"].tolist(), m.matches["offset"].tolist(), m.matches["match_length"].tolist()))
        want = sorted(zip(a.matches["string_idx"].tolist(), a.matches["offset"].tolist(), a.matches["match_length"].tolist()))
        assert got == want
        for k in range(1, shards):
            assert sum(1 for (_, o, l) in want if o < k * per < o + l) >= 1, "no match straddles boundary %d" % k
        # fast mode and a tiny block (not worth sharding: the first device alone) go the same way
        assert _canon(multi.scan_host(data[: 1 << 20], flags=1).matches) == _canon(single.scan_host(data[: 1 << 20], flags=1).matches)
        assert _canon(multi.scan_host(data[:100]).matches) == _canon(single.scan_host(data[:100]).matches)
        assert len(multi.scan_host(b"").matches) == 0
    finally:
        multi.close()
        single.close()


def test_sharded_engine_verifier_error_and_cap():
    """what stays global in a sharded scan: the earliest failing candidate (ERROR_TOO_MANY_RE_FIBERS) and the
    1,000,000-match cap with its TOO_MANY callback"""
    from yara_b200.api import Engine, MultiEngine
    rb = yrh.b200().compile('rule t { strings: $a = "aa" $b = /a1dx((.?)+yy0)/ condition: any of them }')
    image = rb.flatten()
    ndev = Engine.device_count()
    multi, single = MultiEngine(image, [g % ndev for g in range(4)]), Engine(image, 0)
    try:
        data = np.full(6 << 20, ord("a"), dtype=np.uint8)
        m, s = multi.scan_host(data), single.scan_host(data)
        assert m.rc == 0 and s.rc == 0 and m.too_many == s.too_many == [0]
        assert _canon(m.matches) == _canon(s.matches) and len(m.matches) == 1_000_000
        data2 = np.frombuffer(b"." * (5 << 20) + b"zza1dx" + b"." * (1 << 20), dtype=np.uint8)
        m, s = multi.scan_host(data2), single.scan_host(data2)
        assert (m.rc, m.error, m.error_string) == (s.rc, s.error, s.error_string) and s.error == 46
    finally:
        multi.close()
        single.close()


CACHE_CODE = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from yara_b200 import synth
from yara_b200.api import Engine, LibYara, filter_cache_stats
import time
rs = synth.ruleset(3, 4000)
ly = LibYara()
yarc = sys.argv[1]
import os
if os.path.exists(yarc):
    rules = ly.load(yarc)
else:
    rules = ly.compile(rs.source); ly.save(rules, yarc)
image = ly.flatten(rules)
t0 = time.time(); eng = Engine(image, 0); t1 = time.time() - t0
eng2 = Engine(image, 0)
data = synth.corpus(rs, 1 << 22, 2048)
r = eng.scan_host(data)
assert r.rc == 0 and r.error == 0 and r.stats["scan_kernel"] == 2
import hashlib
# the lists of chained strings' heads come in completion order of their chains: compare string-major (stable)
m = r.matches[np.argsort(r.matches["string_idx"], kind="stable")]
print("CACHE", filter_cache_stats(), hashlib.sha1(m.tobytes()).hexdigest(), len(m), round(t1, 3))
"""


def test_filter_tables_come_from_the_disk_cache_in_a_second_process(tmp_path):
    """rules saved with yr_rules_save and loaded with yr_rules_load (arena.c:543/627) in a new
    process find their prefilter tables in the engine's cache keyed by the rules image -- no rebuild -- and the
    scan results are the same."""
    import subprocess, sys
    env = dict(os.environ, YR_B200_CACHE_DIR=str(tmp_path / "cache"))
    outs = []
    for _ in range(2):
        p = subprocess.run([sys.executable, "-c", CACHE_CODE % (yrh.ROOT, os.path.join(yrh.ROOT, "tests")), str(tmp_path / "r.yarc")],
                           env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0 and "CACHE" in p.stdout, (p.stdout + p.stderr)[-3000:]
        outs.append([l for l in p.stdout.splitlines() if l.startswith("CACHE")][0])
    first, second = outs
    assert "'built': 1" in first and "'registry': 1" in first          # built once, the second Engine took it from the registry
    assert "'built': 0" in second and "'disk': 1" in second            # new process, rules from the .yarc: tables from disk
    assert first.split()[-3:-1] == second.split()[-3:-1]               # same matches


DEDUP_CODE = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import cases, yrh
rule, data = cases.CASES[-1]
a = yrh.ref().compile(rule).scan(data)
b = yrh.b200().compile(rule).scan(data)
assert a.rc == b.rc == 0 and a.matches.tobytes() == b.matches.tobytes() and len(a.matches) >= 4, (a.rc, b.rc, len(a.matches))
print("DEDUP ok", len(a.matches))
"""


def test_dedup_moves_the_chain_heads_records_forward():
    """The resolve copies finals[0, n - chain parts) to the host while the chain replay still runs; when plain records
    were deduplicated (scan.c:311-326) the chain heads' lists begin inside that range and are fetched again.  The last
    known-answer case is built to take that path: check that it really does (trace line) and that the result is the
    reference's."""
    import subprocess, sys
    env = dict(os.environ, YRB_TRACE="1")
    p = subprocess.run([sys.executable, "-c", DEDUP_CODE % (yrh.ROOT, os.path.join(yrh.ROOT, "tests"))],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DEDUP ok" in p.stdout, (p.stdout + p.stderr)[-3000:]
    assert "plain records deduplicated" in p.stderr, p.stderr[-3000:]


SHORT_ATOM_RULES = r"""
rule short_atoms {
 strings:
  $mz = "MZ"
  $pk = { 50 4B 03 }
  $q = "q" fullword
  $w = "Zz" wide
  $re1 = /k[0-9]{2,4}!/
  $re2 = /ab\d+c/
  $hx = { 7F ?? 45 [2-6] 46 }
 condition:
  any of them
}
"""


@pytest.mark.parametrize("extra", [0, 1500])
def test_short_atoms_go_through_the_filter_kernel(extra):
    """Rulesets with 1-, 2- and 3-byte atoms (round 1: walk kernel only) run on the filter kernel's short-atom variant:
    an exact 2-byte bitmap probed per position, flagged positions walked from the root (ac_scan.cu, variant 6).  Alone
    and mixed into a 1,500-string synthetic ruleset; data with low-entropy stretches so that the short atoms fire a lot;
    compared with the reference library, all fields, and with the walk kernel."""
    rs = synth.ruleset(3, extra) if extra else None
    src = SHORT_ATOM_RULES + (rs.source if rs else "")
    from yara_b200 import Engine
    rr, rb = both(src)
    image = rb.flatten()
    eng = Engine(image, 0)
    fi = eng.filter_info()
    assert fi["usable"] == 1 and fi["variant"] == 6 and (fi["has_depth1"] or fi["has_depth2"]), fi
    n = (1 << 23) + 4099
    data = synth.corpus(rs, n, 4096) if rs else synth.rand_bytes(99, 0, n)
    data = np.array(data, dtype=np.uint8)
    rng = np.random.RandomState(5)
    for k in range(64):   # stretches of text-like bytes and of one repeated 2-byte atom
        o = int(rng.randint(0, n - 70000))
        data[o:o + 4096] = np.frombuffer((b"MZqk12! ab123c PK\x03 q Z\x00z\x00 \x7f1E..F " * 128)[:4096], dtype=np.uint8)
    data[n // 2: n // 2 + 3000] = np.frombuffer(b"MZ" * 1500, dtype=np.uint8)
    a = rr.scan(data, cap=1 << 21)
    b = rb.scan(data, cap=1 << 21)
    assert a.rc == 0 and len(a.matches) > 20000
    assert_same(a, b)
    from yara_b200 import api as _api
    for kernel in (_api.KERNEL_FILTER, _api.KERNEL_WALK):
        eng.set_option(_api.OPT_SCAN_KERNEL, kernel)
        r = eng.scan_host(data)
        assert r.rc == 0 and r.error == 0 and r.stats["scan_kernel"] == kernel
        got = sorted(zip(r.matches["string_idx"].tolist(), r.matches["offset"].tolist(), r.matches["match_length"].tolist()))
        want = sorted(zip(a.matches["string_idx"].tolist(), a.matches["offset"].tolist(), a.matches["match_length"].tolist()))
        assert got == want
