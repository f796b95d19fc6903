"""GPU tests of the C ABI itself (include/yara_b200.h, through the ctypes
bindings): device-resident scans, byte-range shards with read margins resolved
globally, the match-data gather, kernel selection, and full-size properties."""
import numpy as np
import pytest

import yrh
from yara_b200 import DeviceBuffer, Engine, shard, synth
from yara_b200 import api

pytestmark = pytest.mark.gpu


def as_tuples(m):
    """records in canonical (string, offset) order: the engine returns plain strings' lists string-major and the
    lists of chain heads after them; inside one string the order is always the YR_MATCHES list order"""
    t = [(int(x["string_idx"]), int(x["offset"]), int(x["match_length"]), int(x["xor_key"])) for x in m]
    # list order inside a string must already be ascending offset
    last = {}
    for s_, o, _, _ in t:
        assert last.get(s_, -1) < o
        last[s_] = o
    return sorted(t)


@pytest.fixture(scope="module")
def cfg3():
    rs = synth.ruleset(3, 4000)
    rules = yrh.b200().compile(rs.source)
    image = rules.flatten()
    return rs, rules, image


def test_device_scan_equals_oracle_and_host_scan(cfg3):
    rs, rules, image = cfg3
    eng = Engine(image, 0)
    data = synth.corpus(rs, (1 << 22) + 77, 2048)
    want = as_tuples(yrh.oracle_scan(image, data).matches)
    dbuf = DeviceBuffer(0, len(data))
    dbuf.write(0, data)
    for kernel in (api.KERNEL_WALK, api.KERNEL_FILTER):
        eng.set_option(api.OPT_SCAN_KERNEL, kernel)
        rd = eng.scan_device(dbuf)
        rh = eng.scan_host(data)
        assert rd.rc == 0 and rd.error == 0 and rh.rc == 0
        assert rd.stats["scan_kernel"] == kernel
        assert as_tuples(rd.matches) == want
        assert as_tuples(rh.matches) == want
    # match data copy for HBM-resident blocks (scan.c:723-737); it follows the record order of the
    # scanner's LAST scan (the lists of chain heads come out in no fixed order between scans)
    rd = eng.scan_device(dbuf)
    blob, offs = eng.gather_match_data(dbuf, len(rd.matches), 512)
    for i in range(0, len(rd.matches), 7):
        o, l = int(rd.matches["offset"][i]), min(int(rd.matches["match_length"][i]), 512)
        assert bytes(blob[int(offs[i]) : int(offs[i + 1])]) == bytes(data[o : o + l])


def test_device_random_fill_matches_host_generator():
    dbuf = DeviceBuffer(0, 100003)
    for off in (0, 5, 8, 4099):
        dbuf.fill_random(7, off)
        assert (dbuf.read(0, 100003) == synth.rand_bytes(7, off, 100003)).all()


@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_byte_range_shards_equal_single_block(cfg3, nshards):
    """shards with margins + one global resolve == one single-block scan,
    including plants straddling shard boundaries and chain parts in different shards."""
    rs, rules, image = cfg3
    eng = Engine(image, 0)
    size = 3 * (1 << 20) + 4096 * 5
    data = synth.corpus(rs, size, 1024)
    margin = shard.required_margin(image)
    # plant instances across every shard boundary (literal/hex and chained strings)
    import random
    rng = random.Random(5)
    chained = [sp for sp in rs.specs if sp.kind == "hexwild" and "[2" in sp.text or "[3" in sp.text][:50]
    for lo, hi, g0, g1 in shard.plan(size, nshards, margin)[1:]:
        for sp in (rs.specs[1], rs.specs[2], rs.specs[7]) + tuple(chained[:3]):
            inst = np.frombuffer(sp.instance(rng), dtype=np.uint8)
            start = lo - len(inst) // 2
            if start > 0 and start + len(inst) < size:
                data[start : start + len(inst)] = inst
    whole = eng.scan_host(data)
    assert whole.rc == 0 and whole.error == 0
    sharded = shard.scan_sharded(eng, data, nshards, margin)
    assert sharded.rc == 0 and sharded.error == 0
    assert as_tuples(sharded.matches) == as_tuples(whole.matches)
    assert as_tuples(whole.matches) == as_tuples(yrh.oracle_scan(image, data).matches)


def test_shard_without_margin_is_refused(cfg3):
    rs, rules, image = cfg3
    eng = Engine(image, 0)
    data = synth.corpus(rs, 1 << 20, 1024)
    half = 1 << 19
    # second shard handed over with no look-back bytes at all
    r = eng.scan_host(data[half:], data_offset=half, block_size=len(data), own=(half, len(data)), defer=True)
    assert r.rc == 1005  # YRB_ERROR_MARGIN


def test_full_size_properties():
    """1 GiB (BASELINE configs[1] size): every planted instance is reported at its
    planted offset; scanning twice is idempotent; walk and filter kernels agree."""
    rs = synth.ruleset(2, 1000)
    rules = yrh.b200().compile(rs.source)
    image = rules.flatten()
    eng = Engine(image, 0)
    size = 1 << 30
    dbuf = DeviceBuffer(0, size)
    dbuf.fill_random(7, 0)
    pl = synth.plants(rs, size, 65536, 11)
    assert dbuf.scatter(pl) == len(pl)
    r1 = eng.scan_device(dbuf)
    assert r1.rc == 0 and r1.error == 0
    got = set(zip(r1.matches["offset"].tolist(), r1.matches["match_length"].tolist()))
    for off, inst in pl:
        assert (off, len(inst)) in got
    r2 = eng.scan_device(dbuf)
    assert r1.matches.tobytes() == r2.matches.tobytes()
    eng.set_option(api.OPT_SCAN_KERNEL, api.KERNEL_WALK)
    r3 = eng.scan_device(dbuf)
    assert r1.matches.tobytes() == r3.matches.tobytes()
    # sortedness / uniqueness of every YR_MATCHES list
    key = r1.matches["string_idx"].astype(np.int64) * (1 << 40) + r1.matches["offset"]
    assert (np.diff(key) > 0).all()


def test_scan_timeout_takes_effect_between_h2d_slices():
    """scanner.c:74-81: the reference checks its stopwatch every 4096 bytes and returns ERROR_SCAN_TIMEOUT (26).
    The engine checks between H2D slices (at most four queued ahead): a 2 GiB host block with a budget far below
    its copy time times out MID-scan -- well before the whole block has been copied -- and the same scan with a
    generous budget succeeds; through the drop-in (yr_scanner_scan_mem) likewise."""
    import time
    rs = synth.ruleset(3, 2000)
    rb = yrh.b200().compile(rs.source)
    eng = Engine(rb.flatten(), 0)
    data = synth.rand_bytes(7, 0, 1 << 31)
    t0 = time.perf_counter(); ok = eng.scan_host(data); full = time.perf_counter() - t0
    assert ok.rc == 0
    t0 = time.perf_counter(); r = eng.scan_host(data, timeout_ns=2_000_000); cut = time.perf_counter() - t0
    assert r.rc == 26                       # ERROR_SCAN_TIMEOUT
    assert cut < full / 2, (cut, full)      # given up long before the end of the block
    assert eng.scan_host(data, timeout_ns=600_000_000_000).rc == 0
    assert eng.scan_host(data[: 1 << 20], timeout_ns=2_000_000_000).rc == 0
    # the unchanged libyara API: scanner->timeout (set in ns by the harness; yr_scanner_set_timeout takes seconds)
    assert rb.scan_timeout_ns(data, 2_000_000) == 26
    assert rb.scan_timeout_ns(data, 600_000_000_000) == 0
    eng.close()
