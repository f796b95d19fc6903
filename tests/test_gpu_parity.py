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
