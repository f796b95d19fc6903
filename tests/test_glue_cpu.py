"""CPU tests of the libyara glue (yara_b200/host/yr_b200_glue.c + the hook lines), end to end through the
unchanged libyara API: tests/harness/libyara_b200cpu.so is the drop-in's own objects linked against a
TEST stand-in for the CUDA engine (tests/harness/fake_engine.cpp) that replays the product's verify /
resolve device code on the CPU.  What the GPU tests check on a B200 -- YR_MATCH materialisation, cross-block
chains (seeds), parked file-batch results, the pre-filter's note site -- is checked here against the reference
library without a GPU.  The product itself has no CPU path (tests/test_abi_cpu.py::test_no_cpu_fallback)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import fuzzgen
import yrh
from cases import CASES
from yara_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = os.path.exists(yrh.b200cpu_path()) and os.path.exists(yrh.ref_path())
pytestmark = pytest.mark.skipif(not HAVE, reason="tests/harness/libyrh_b200cpu.so or oracle/_ref not built")


def same(a, b, what=""):
    assert a.rc == b.rc, (what, a.rc, b.rc)
    if a.rc == 0:
        assert a.matches.tobytes() == b.matches.tobytes(), what
        assert a.rule_bits.tobytes() == b.rule_bits.tobytes(), what
    assert a.too_many_calls == b.too_many_calls, what


def capped(ref_result):
    return False      # the stand-in replays the second pass of the 1M cap too


def run_cases():
    R, B = yrh.ref(), yrh.b200cpu()
    n = 0
    for idx, (rule, data) in enumerate(CASES):
        rr, rb = R.compile(rule), B.compile(rule)
        for flags in (0, 1):
            a = rr.scan(data, flags=flags, cap=1 << 20)
            if capped(a):
                continue
            same(a, rb.scan(data, flags=flags, cap=1 << 20), "case %d flags %d" % (idx, flags))
            n += 1
        for flags in (0, 1):     # fast mode over several blocks: chains completed by a later block's tail still count
            a = rr.scan_blocks(data, 1024, 256, flags=flags, cap=1 << 20)
            if not capped(a):
                same(a, rb.scan_blocks(data, 1024, 256, flags=flags, cap=1 << 20), "case %d blocks flags %d" % (idx, flags))
                n += 1
    return n


def run_fuzz(seeds):
    R, B = yrh.ref(), yrh.b200cpu()
    for seed in seeds:
        text, insts, n = fuzzgen.compile_filtered(R, seed, 50)
        rr, rb = R.compile(text), B.compile(text)
        data = np.frombuffer(bytes(fuzzgen.gen_data(seed, insts, (1 << 14) + 13 * seed)), dtype=np.uint8)
        for flags in (0, 1):
            a = rr.scan(data, flags=flags, cap=1 << 20)
            if not capped(a):
                same(a, rb.scan(data, flags=flags, cap=1 << 20), "seed %d flags %d" % (seed, flags))
        for flags in (0, 1):
            a = rr.scan_blocks(data, 1024, 256, flags=flags, cap=1 << 20)          # chains whose parts lie in different blocks: seeds
            if not capped(a):
                same(a, rb.scan_blocks(data, 1024, 256, flags=flags, cap=1 << 20), "seed %d blocks flags %d" % (seed, flags))
        # the same data as a parked file batch
        cuts = sorted(set([0, len(data)] + [(seed * 7919 + k * 10007) % len(data) for k in range(6)]))
        items = [np.ascontiguousarray(data[a_:b_]) for a_, b_ in zip(cuts, cuts[1:])] + [data]
        want = [rr.scan(it, cap=1 << 20) for it in items]
        with rb.batch(items) as bt:
            assert bt.rc == 0
            for k, it in enumerate(items):
                if not capped(want[k]):
                    same(want[k], rb.scan(it, cap=1 << 20), "seed %d batch item %d" % (seed, k))


def run_file_batch():
    rs = synth.ruleset(4, 1500)
    R, B = yrh.ref(), yrh.b200cpu()
    rr, rb = R.compile(rs.source), B.compile(rs.source)
    files = [synth.corpus(rs, 9000 + 700 * k, 2048, seed=k, plant_seed=30 + k) for k in range(24)]
    want = [rr.scan(f) for f in files]
    h0 = yrh.batch_hits(B)
    with rb.batch(files) as bt:
        assert bt.rc == 0 and bt.n_rescans == 0
        for rounds in range(2):
            for k, f in enumerate(files):
                same(want[k], rb.scan(f), "file %d" % k)
        assert yrh.batch_hits(B) - h0 == 2 * len(files)     # every one of them came from the parked results
    for k, f in enumerate(files[:6]):      # and without the batch
        same(want[k], rb.scan(f), "file %d unbatched" % k)
    assert yrh.batch_hits(B) - h0 == 2 * len(files)


def test_known_answer_cases_through_the_glue():
    assert run_cases() > 100


def test_fuzz_rulesets_through_the_glue():
    run_fuzz(range(8))


def test_file_batch_through_the_glue():
    run_file_batch()


def test_the_same_with_the_other_prefilter_setting():
    """$YR_B200_RULE_PREFILTER is read once per process (on unless "0"): run the checks above once more in a fresh
    process with the opposite setting"""
    env = dict(os.environ)
    on_here = not env.get("YR_B200_RULE_PREFILTER", "").startswith("0")
    env["YR_B200_RULE_PREFILTER"] = "0" if on_here else "1"
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_glue_cpu as t; "
            "t.run_cases(); t.run_fuzz(range(4)); t.run_file_batch(); print('GLUE_OK')" % (ROOT, os.path.join(ROOT, "tests")))
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "GLUE_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]


@pytest.mark.parametrize("prefilter", ["0", "1"])
@pytest.mark.parametrize("name", ["test-rules_b200cpu", "test-api_b200cpu", "test-async_b200cpu"])
def test_reference_programs_through_the_glue(name, prefilter):
    """the reference's own test programs (test-rules: ~1,500 assertions x 3 passes incl. 1024-byte blocks; test-api
    incl. test_too_many_matches), compiled unmodified and linked against the stand-in build of the drop-in"""
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip("%s not built" % name)
    env = dict(os.environ, TOP_SRCDIR=os.path.join(ROOT, "oracle", "_ref", "top"))
    env.pop("YR_B200_RULE_PREFILTER", None)
    if prefilter:
        env["YR_B200_RULE_PREFILTER"] = prefilter
    p = subprocess.run([exe], cwd=os.path.join(ROOT, "oracle", "_ref"), env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]


def test_parked_result_is_not_used_for_the_first_block_of_a_multi_block_scan():
    """a parked batch item carries the COMPLETE result of a one-block scan (leftover unconfirmed chain parts are
    dropped, as at the end of a scan); if the same bytes come in as the first block of a multi-block scan, the
    chain parts they leave behind must survive into the next block, so the parked result must not be used"""
    rule = "rule t { strings: $a = { 11 22 33 44 [300-400] 55 66 77 88 } condition: $a }"
    R, B = yrh.ref(), yrh.b200cpu()
    rr, rb = R.compile(rule), B.compile(rule)
    # the reference compares raw block-relative offsets across blocks (scan.c:474-488): head at offset 100 of block 0,
    # tail at offset 450 of block 1
    data = np.frombuffer(b"." * 100 + bytes.fromhex("11223344") + b"." * (1024 + 450 - 104) + bytes.fromhex("55667788") + b"." * 300,
                         dtype=np.uint8).copy()
    want = rr.scan_blocks(data, 1024, 0)
    assert len(want.matches) == 1                      # head in block 0, tail in block 1
    first_block = data[:1024]                          # same address and size as block 0 of the scan below
    with rb.batch([first_block]) as bt:
        assert bt.rc == 0
        same(rr.scan(first_block), rb.scan(first_block), "the parked item itself")
        same(want, rb.scan_blocks(data, 1024, 0), "multi-block scan over a parked first block")


COALESCE_CODE = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import yrh
from yara_b200 import synth
rs = synth.ruleset(4, 1500)
R, B = yrh.ref(), yrh.b200cpu()
rr, rb = R.compile(rs.source), B.compile(rs.source)
files = [synth.corpus(rs, 9000 + 531 * k, 2048, seed=k, plant_seed=40 + k) for k in range(64)]
files.append(np.zeros(0, np.uint8))
want = [rr.scan(f).key() for f in files]
for rounds in range(3):
    with ThreadPoolExecutor(8) as ex:
        got = list(ex.map(lambda f: rb.scan(f).key(), files))
    assert got == want, "round %%d" %% rounds
# fast mode and multi-block scans never go through a shared batch
assert rb.scan(files[3], flags=1).key() == rr.scan(files[3], flags=1).key()
assert rb.scan_blocks(files[5], 1024, 256).key() == rr.scan_blocks(files[5], 1024, 256).key()
b, s = C.c_uint64(), C.c_uint64()
B.lib.yrh_coalesce_stats(C.byref(b), C.byref(s))
print("COALESCE_OK batches %%d scans %%d" %% (b.value, s.value))
"""


@pytest.mark.parametrize("n", ["8", "2"])
def test_request_coalescing(n):
    """$YR_B200_COALESCE=N: concurrent one-block scans of the same rules share one yrb_scan_batch call; results
    stay those of the reference.  (Read once per process, hence the subprocess.)"""
    env = dict(os.environ, YR_B200_COALESCE=n)
    p = subprocess.run([sys.executable, "-c", COALESCE_CODE % (ROOT, os.path.join(ROOT, "tests"))], env=env, capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0 and "COALESCE_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("COALESCE_OK")][0].split()
    batches, scans = int(line[2]), int(line[4])
    assert scans == 3 * 64                      # every non-empty one-block scan went through a shared batch
    assert batches < scans                      # ... and some batches really held more than one scan


def test_concurrent_scans_through_the_glue():
    """yr_rules_scan_mem from several threads at once over one YR_RULES (limits.h YR_MAX_THREADS), no batch parked:
    every thread goes through the engine on its own pooled scanner"""
    from concurrent.futures import ThreadPoolExecutor
    rs = synth.ruleset(4, 1500)
    R, B = yrh.ref(), yrh.b200cpu()
    rr, rb = R.compile(rs.source), B.compile(rs.source)
    files = [synth.corpus(rs, 20000 + 913 * k, 1024, seed=k, plant_seed=20 + k) for k in range(32)]
    want = [rr.scan(f).key() for f in files]
    for rounds in range(3):
        with ThreadPoolExecutor(8) as ex:
            got = list(ex.map(lambda f: rb.scan(f).key(), files))
        assert got == want
    secs, rc = rb.scan_mt(files[0], 8)
    assert rc == 0


MULTI_CODE = r"""
import ctypes as C, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import yrh
from yara_b200 import synth
rs = synth.ruleset(3, 1200)
R, B = yrh.ref(), yrh.b200cpu()
rr, rb = R.compile(rs.source), B.compile(rs.source)
eng = C.CDLL(yrh.b200cpu_path().replace("libyrh_b200cpu.so", "libyrb200cpu.so"))
big = synth.corpus(rs, (1 << 21) + 4321, 2048)
small = synth.corpus(rs, 60000, 2048, seed=3)
before = eng.yrb_test_multi_scans()
assert rb.scan(big, cap=1 << 18).key() == rr.scan(big, cap=1 << 18).key()            # >= 1 MiB, one block: sharded path
mid = eng.yrb_test_multi_scans()
assert rb.scan(small).key() == rr.scan(small).key()                                  # below the threshold: one device
assert rb.scan_blocks(big, 1 << 20, 4096, cap=1 << 18).key() == rr.scan_blocks(big, 1 << 20, 4096, cap=1 << 18).key()
after = eng.yrb_test_multi_scans()
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(4) as ex:                                                    # the sharded engine is shared: one at a time
    got = list(ex.map(lambda k: rb.scan(big, cap=1 << 18).key(), range(8)))
assert all(g == rr.scan(big, cap=1 << 18).key() for g in got)
print("MULTI_OK", mid - before, after - mid)
"""


def test_large_single_block_scans_go_to_the_sharded_engine():
    """$YR_B200_DEVICES lists several GPUs: a block of at least $YR_B200_MULTI_MIN_MIB that is the only block of
    its scan goes through yrb_multi_scan_host; smaller blocks stay on one device.  Blocks of a
    multi-block scan carry chain state (seeds) from block to block; from the second block on -- as soon as
    there is something to carry -- they are not sharded.  Results stay those of the reference."""
    env = dict(os.environ, YR_B200_DEVICES="0,1,2,3", YR_B200_MULTI_MIN_MIB="1")
    p = subprocess.run([sys.executable, "-c", MULTI_CODE % (ROOT, os.path.join(ROOT, "tests"))], env=env, capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0 and "MULTI_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("MULTI_OK")][0].split()
    assert int(line[1]) == 1          # the big one-block scan was sharded


BLOCKS_CODE = r"""
import ctypes as C, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import yrh
from yara_b200 import synth
rs = synth.ruleset(2, 300)                      # literals only: no chained strings
rc = synth.ruleset(3, 600)                      # with chained hex strings
R, B = yrh.ref(), yrh.b200cpu()
eng = C.CDLL(yrh.b200cpu_path().replace("libyrh_b200cpu.so", "libyrb200cpu.so"))
out = []
for r_ in (rs, rc):
    rr, rb = R.compile(r_.source), B.compile(r_.source)
    data = synth.corpus(r_, 48 * 1024 + 100, 1024)
    h0, b0 = eng.yrb_test_host_scans(), eng.yrb_test_batch_calls()
    got = rb.scan_blocks(data, 1024, 256, cap=1 << 18)
    h1, b1 = eng.yrb_test_host_scans(), eng.yrb_test_batch_calls()
    want = rr.scan_blocks(data, 1024, 256, cap=1 << 18)
    assert got.key() == want.key() and len(want.matches) > 20
    out += [h1 - h0, b1 - b0]
print("BLOCKS_OK", *out)
"""


def test_small_blocks_of_a_multi_block_scan_are_batched():
    """the 1 KiB blocks of a multi-block scan (63 of them here) are deferred by
    yr_b200_scan_mem_block and scanned as ONE engine batch behind the block loop (yr_b200_blocks_flush);
    blocks in which a chain part matched are rescanned one by one with the chain state as seeds.  Results are the
    reference's; with $YR_B200_BLOCK_BATCH=0 every block is an engine call of its own."""
    code = BLOCKS_CODE % (ROOT, os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "BLOCKS_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
    h_plain, b_plain, h_chain, b_chain = [int(x) for x in [l for l in p.stdout.splitlines() if l.startswith("BLOCKS_OK")][0].split()[1:]]
    assert h_plain == 0 and b_plain == 1          # no chains: every block answered by the one batch
    assert b_chain == 1 and 0 < h_chain < 63      # only the blocks with chain-part matches were scanned on their own
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=dict(os.environ, YR_B200_BLOCK_BATCH="0"))
    assert p.returncode == 0 and "BLOCKS_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
    h_plain, b_plain, h_chain, b_chain = [int(x) for x in [l for l in p.stdout.splitlines() if l.startswith("BLOCKS_OK")][0].split()[1:]]
    assert b_plain == 0 and b_chain == 0 and h_plain >= 63 and h_chain >= 63


PACKED_CODE = r"""
import ctypes as C, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import yrh
R, B = yrh.ref(), yrh.b200cpu()
eng = C.CDLL(yrh.b200cpu_path().replace("libyrh_b200cpu.so", "libyrb200cpu.so"))
rule = 'rule t { strings: $a = "abcd" $b = /ef[0-9]{1,600}gh/ $c = "zz" private condition: any of them }'
data = np.frombuffer(b"abcd ef123gh " * 9000 + b"ef" + b"7" * 550 + b"gh" + b"abcd" * 300, dtype=np.uint8)
rr, rb = R.compile(rule), B.compile(rule)
c0 = eng.yrb_test_match_data_calls()
a, b = rr.scan(data, cap=1 << 18), rb.scan(data, cap=1 << 18)
c1 = eng.yrb_test_match_data_calls()
assert a.rc == b.rc == 0 and len(a.matches) > 18000 and a.matches.tobytes() == b.matches.tobytes()
small = data[:4000]
assert rr.scan(small).matches.tobytes() == rb.scan(small).matches.tobytes()
c2 = eng.yrb_test_match_data_calls()
print("PACKED_OK", c1 - c0, c2 - c1)
"""


def test_large_results_take_their_match_data_packed_from_the_engine():
    """materialise() copies min(match_length, max_match_data) bytes per match (scan.c:723-737).  For results of at
    least 8192 matches the glue asks the engine for them packed (yrb_last_match_data: gathered from the resident copy
    of the block) instead of reading the host block at random; every YR_MATCH field incl. the data hash must stay the
    reference's, also for matches longer than max_match_data (512)."""
    for env_extra, want in (({}, 1), ({"YR_B200_GATHER_MATCH_DATA": "0"}, 0)):
        p = subprocess.run([sys.executable, "-c", PACKED_CODE % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True,
                           timeout=900, env=dict(os.environ, **env_extra))
        assert p.returncode == 0 and "PACKED_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]
        big, small = [int(x) for x in [l for l in p.stdout.splitlines() if l.startswith("PACKED_OK")][0].split()[1:]]
        assert big == want and small == 0


def run_long_chains(B, n=600):
    """tests/long_chains.py: hundreds of rules sharing one string -> one AC state with hundreds of YR_AC_MATCH
    entries.  Whole block in both modes, 1024-byte blocks (the shared chain holds heads and tails of chained
    strings: seeds across blocks), and the same bytes as a parked file batch, against the reference library."""
    import long_chains
    R = yrh.ref()
    text = long_chains.ruleset(n)
    rr, rb = R.compile(text), B.compile(text)
    data = long_chains.corpus()
    total = 0
    for flags in (0, 1):
        a = rr.scan(data, flags=flags, cap=1 << 20)
        same(a, rb.scan(data, flags=flags, cap=1 << 20), "long chains flags %d" % flags)
        total += len(a.matches)
    same(rr.scan_blocks(data, 1024, 256, cap=1 << 20), rb.scan_blocks(data, 1024, 256, cap=1 << 20), "long chains blocks")
    items = [np.ascontiguousarray(data[a_:b_]) for a_, b_ in ((0, 5000), (5000, 5003), (5003, 20000), (20000, len(data)))] + [data]
    want = [rr.scan(it, cap=1 << 20) for it in items]
    with rb.batch(items) as bt:
        assert bt.rc == 0
        for k, it in enumerate(items):
            same(want[k], rb.scan(it, cap=1 << 20), "long chains batch item %d" % k)
    return total


def test_long_match_chains_through_the_glue():
    assert run_long_chains(yrh.b200cpu()) > 6000


def run_scanner_reuse_after_timed_out_scan(B):
    """A scanner object is reused after a multi-block scan that ended early (scanner.c:540-552 cleans up behind every
    scan, whatever its result): blocks the drop-in had deferred when the first scan ran out of time must not be
    scanned as part of the next one.  The iterator of the first scan turns slow after a few 4 KiB blocks."""
    rs = synth.ruleset(3, 300)
    R = yrh.ref()
    rr, rb = R.compile(rs.source), B.compile(rs.source)
    data1 = synth.corpus(rs, 1 << 18, 512, seed=3, plant_seed=4)
    data2 = synth.corpus(rs, 1 << 17, 2048, seed=5, plant_seed=6)
    want = rr.scan_blocks(data2, 4096, 0)
    assert want.rc == 0 and len(want.matches) > 30
    for r in (rr, rb):
        rc1, got = r.scan_blocks_after_timed_out_scan(data1, 5 * 4096, 20_000_000, data2, 4096)
        assert rc1 == 26                   # ERROR_SCAN_TIMEOUT
        same(want, got, "second scan of a reused scanner")


def test_scanner_reuse_after_a_scan_that_ended_early():
    run_scanner_reuse_after_timed_out_scan(yrh.b200cpu())


def run_scanner_sequence(B, seeds=range(6)):
    """One long-lived YR_SCANNER, many scans of different kinds one after the other -- whole buffers, 1 KiB and 4 KiB
    blocks, fast mode switched on and off in between, an empty buffer -- as a consumer of the yr_scanner_* API does.
    Every scan's matches and rule verdicts against the reference library driven the same way."""
    R = yrh.ref()
    for seed in seeds:
        text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
        rr, rb = R.compile(text), B.compile(text)
        datas = [np.frombuffer(bytes(fuzzgen.gen_data(seed + 100 * k, insts, (1 << 13) + 97 * k)), dtype=np.uint8) for k in range(4)]
        scans = [(datas[0], 0, 0), (datas[1], 1024, 0), (datas[0], 0, 1), (datas[2], 4096, 0), (datas[1], 0, 0),
                 (np.zeros(0, dtype=np.uint8), 0, 0), (datas[3], 1024, 1), (datas[3], 1024, 0), (datas[2], 0, 0), (datas[0], 0, 0)]
        want = rr.scanner_sequence(scans)
        got = rb.scanner_sequence(scans)
        for k, (a, b) in enumerate(zip(want, got)):
            assert a[0] == b[0], ("seed %d scan %d rc" % (seed, k), a[0], b[0])
            if a[0] == 0:
                assert a[1].tobytes() == b[1].tobytes(), "seed %d scan %d matches" % (seed, k)
                assert a[2].tobytes() == b[2].tobytes(), "seed %d scan %d rules" % (seed, k)
        assert want[0][1].tobytes() == want[9][1].tobytes()      # the same scan again at the end: same result


def test_one_scanner_many_scans():
    run_scanner_sequence(yrh.b200cpu())


def region_lists(size, rng):
    """block lists a process-memory or custom iterator may produce: ascending with gaps, descending, overlapping,
    repeated, empty blocks, bases unrelated to the position in the buffer"""
    step = 1500
    asc = [(o, min(step, size - o), o) for o in range(0, size, step + 700)]
    yield "ascending with gaps", asc
    yield "descending", asc[::-1]
    over = [(o, min(2048, size - o), o) for o in range(0, size - 1, 1024)]
    yield "overlapping", over
    yield "every block twice", [r for r in over for _ in (0, 1)]
    yield "with empty blocks", [x for r in asc for x in (r, (r[0], 0, r[2]))]
    sh = list(over)
    rng.shuffle(sh)
    yield "shuffled", sh
    yield "foreign bases", [(o, n, 0x7F0000000000 + 3 * o) for o, n, _ in asc]
    yield "all at base 0", [(o, n, 0) for o, n, _ in asc]
    mixed, o = [], 0
    while o < size:
        n = min(rng.choice([300, 900, 1500, 1700, 2500, 4000]), size - o)     # around $YR_B200_BLOCK_BATCH of the test below
        mixed.append((o, n, o))
        o += n - min(n // 2, 128)
    yield "mixed sizes", mixed


def run_regions(B, seeds=range(5)):
    import random
    R = yrh.ref()
    for seed in seeds:
        text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
        rr, rb = R.compile(text), B.compile(text)
        data = np.frombuffer(bytes(fuzzgen.gen_data(seed + 7, insts, (1 << 14) + 11 * seed)), dtype=np.uint8)
        for what, regions in region_lists(len(data), random.Random(seed)):
            for flags in (0, 1):
                same(rr.scan_regions(data, regions, flags), rb.scan_regions(data, regions, flags), "seed %d %s flags %d" % (seed, what, flags))


def test_block_lists_of_any_shape():
    run_regions(yrh.b200cpu())


def test_errors_of_strings_settled_by_earlier_blocks_do_not_count():
    """Fast mode over several blocks, fuzz seeds 1325 and 1563 (found by scripts/glue_sweep.py): a SINGLE_MATCH string
    matches in an early block; in a later block one of its candidates would overflow the fiber pool
    (ERROR_TOO_MANY_RE_FIBERS) -- but the reference no longer verifies that string (scan.c:1021-1023) and finishes
    the scan.  The glue asks the engine again with the string named in yrb_scan_params.skip_strings."""
    run_fuzz([1325, 1563])


def test_block_lists_with_a_small_deferral_limit():
    """the same shapes with $YR_B200_BLOCK_BATCH=1600 (read once per process): blocks above it are scanned at once and
    flush the deferred ones before them, so deferred and immediate blocks alternate"""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import yrh, test_glue_cpu as t; "
            "t.run_regions(yrh.b200cpu(), range(3)); print('REGIONS_OK')" % (ROOT, os.path.join(ROOT, "tests")))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=dict(os.environ, YR_B200_BLOCK_BATCH="1600"))
    assert p.returncode == 0 and "REGIONS_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]


MMD_CODE = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import fuzzgen, yrh, test_glue_cpu as t
R, B = yrh.ref(), yrh.b200cpu()
for mmd in (0, 7, 4096):
    assert R.lib.yrh_set_max_match_data(mmd) == 0 and B.lib.yrh_set_max_match_data(mmd) == 0
    t.run_fuzz(range(3))
    # a result large enough for the packed match data path (8192+ matches)
    rule = 'rule r { strings: $a = /ab+c?/ $b = "abbbbbbbbbbbbbbbbbbbb" condition: any of them }'
    data = np.frombuffer((b"abbbbbbbbbbbbbbbbbbbbbbbbc--" * 6000), dtype=np.uint8)
    a, b = R.compile(rule).scan(data, cap=1 << 20), B.compile(rule).scan(data, cap=1 << 20)
    assert len(a.matches) >= 12000
    t.same(a, b, "max_match_data %%d" %% mmd)
print("MMD_OK")
"""


def test_other_settings_of_max_match_data():
    """YR_CONFIG_MAX_MATCH_DATA (libyara.c:306; default 512) is process-wide: 0 (no data kept), a few bytes, 4096 --
    data_length and the copied bytes (hashed by the harness) follow it, in materialise() and in the packed path."""
    p = subprocess.run([sys.executable, "-c", MMD_CODE % (ROOT, os.path.join(ROOT, "tests"))], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "MMD_OK" in p.stdout, (p.stdout + p.stderr)[-3000:]


def run_cap_across_blocks(B):
    """YR_MAX_STRING_MATCHES (1,000,000) reached across the blocks of one scan: inside the first block, in the middle of
    a later block (the list already holds what earlier blocks found), with the callback answering CONTINUE (the string
    is disabled for the rest of the scan, later blocks add nothing) and ABORT (ERROR_TOO_MANY_MATCHES).  A second string
    keeps matching throughout."""
    R = yrh.ref()
    rule = 'rule r { strings: $a = "aa" $b = "aab" condition: any of them }'
    rr, rb = R.compile(rule), B.compile(rule)
    data = np.full(2_300_000, ord("a"), dtype=np.uint8)
    data[::50_000] = ord("b")
    for block_size in (1_200_000, 700_000, 300_000, 200_000):     # the last: small enough to be deferred and batched
        for action in (0, 1):
            a = rr.scan_blocks(data, block_size, 0, too_many_action=action, cap=1 << 21)
            b = rb.scan_blocks(data, block_size, 0, too_many_action=action, cap=1 << 21)
            same(a, b, "cap across blocks of %d, action %d" % (block_size, action))
            assert a.too_many_calls == 1 and (a.rc == 0) == (action == 0)


def test_match_cap_across_blocks():
    run_cap_across_blocks(yrh.b200cpu())


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_block_geometries(seed):
    """other block sizes and overlaps than the reference's own test pass uses -- tiny blocks, overlaps of all but one
    byte (every match is found many times over: dedup by base + offset), no overlap -- in both modes"""
    R, B = yrh.ref(), yrh.b200cpu()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
    rr, rb = R.compile(text), B.compile(text)
    data = np.frombuffer(bytes(fuzzgen.gen_data(seed + 900, insts, 3000 + 17 * seed)), dtype=np.uint8)
    for bs, ov in ((64, 16), (100, 99), (257, 0), (4096, 0), (1000, 500), (33, 32)):
        for flags in (0, 1):
            same(rr.scan_blocks(data, bs, ov, flags=flags, cap=1 << 20), rb.scan_blocks(data, bs, ov, flags=flags, cap=1 << 20),
                 "seed %d blocks %d/%d flags %d" % (seed, bs, ov, flags))


def test_a_fuzz_ruleset_repeated_66_times():
    """every rule of fuzz ruleset 1 under 66 names: match chains of 264 entries (more than an 8-bit chain position
    holds) over all string types at once -- literals, hex with jumps and chains, regexps -- through the glue in all
    the modes of run_fuzz"""
    import re
    orig = fuzzgen.compile_filtered

    def repeated(R, seed, n):
        text, insts, cnt = orig(R, seed, n)
        return "".join(re.sub(r"\brule (\w+)", lambda m: "rule %s_%d" % (m.group(1), k), text) for k in range(66)), insts, cnt * 66

    fuzzgen.compile_filtered = repeated
    try:
        run_fuzz([1])
    finally:
        fuzzgen.compile_filtered = orig


def test_concurrent_multi_block_scans():
    """eight threads, each with its own scanner, scanning different buffers in 1 KiB and 4 KiB blocks in both modes at
    the same time over ONE YR_RULES (the deferred-block state is per scanner; device scanners come from the shared pool)"""
    from concurrent.futures import ThreadPoolExecutor
    R, B = yrh.ref(), yrh.b200cpu()
    text, insts, n = fuzzgen.compile_filtered(R, 2, 40)
    rr, rb = R.compile(text), B.compile(text)
    datas = [np.frombuffer(bytes(fuzzgen.gen_data(2 + 50 * k, insts, (1 << 14) + 31 * k)), dtype=np.uint8) for k in range(12)]
    jobs = [(k, bs, fl) for k in range(12) for bs in (1024, 4096) for fl in (0, 1)]
    want = {j: rr.scan_blocks(datas[j[0]], j[1], j[1] // 4, flags=j[2], cap=1 << 20) for j in jobs}

    def run(j):
        return j, rb.scan_blocks(datas[j[0]], j[1], j[1] // 4, flags=j[2], cap=1 << 20)

    for rounds in range(2):
        with ThreadPoolExecutor(8) as ex:
            for j, got in ex.map(run, jobs):
                same(want[j], got, "job %s" % (j,))
