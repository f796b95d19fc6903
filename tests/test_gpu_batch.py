"""GPU tests of yrb_scan_batch (include/yara_b200.h): a file batch scanned with one
launch sequence must give, per item, exactly what a scan of that item alone gives --
the reference scans every file from the root state with its own bounds
(cli/yara.c:365-457 -> yr_rules_scan_file -> scanner.c:45-176)."""
import random

import numpy as np
import pytest

import yrh
from yara_b200 import Engine, PinnedBuffer, synth
from yara_b200 import api

pytestmark = pytest.mark.gpu


def tuples(m):
    return [(int(x["string_idx"]), int(x["offset"]), int(x["match_length"]), int(x["xor_key"])) for x in m]


def canon(m):
    return sorted(tuples(m))


@pytest.fixture(scope="module", params=[4, 5])
def setup(request):
    # config 4: literals, hex with jumps (chains), nocase/wide/fullword, regexes; config 5: regexes, xor/base64
    rs = synth.ruleset(request.param, 1500)
    rules = yrh.b200().compile(rs.source)
    image = rules.flatten()
    return rs, image


def make_items(rs, seed=3):
    rng = random.Random(seed)
    sizes = [0, 1, 3, 4, 5, 17, 100, 4096, 65536, 65537, 1 << 20, (1 << 20) + 3, 300000, 0, 70000, 2 * (1 << 20) + 11, 9, 33000]
    sizes += [rng.randrange(8000, 70000) for _ in range(40)]   # small enough that many hold no chained-string plant
    items = []
    for k, n in enumerate(sizes):
        if n >= 4096:
            it = synth.corpus(rs, n, 4096 if n < 200000 else 2048, seed=100 + k, plant_seed=200 + k)
        else:
            it = synth.rand_bytes(100 + k, 0, n).copy()
        items.append(it)
    # instances cut in two by an item boundary must match in neither half; full instances at the very
    # start / end of an item must match
    plain = [sp for sp in rs.specs if sp.kind in ("text", "hexlit", "special", "textmod")][:6]
    for j, sp in enumerate(plain):
        inst = np.frombuffer(sp.instance(rng), dtype=np.uint8)
        a, b = items[8 + j % 4], items[9 + j % 4]
        cut = max(1, len(inst) // 2)
        if len(a) > len(inst) and len(b) > len(inst):
            a[len(a) - cut:] = inst[:cut]
            b[: len(inst) - cut] = inst[cut:]
    inst = np.frombuffer(plain[0].instance(rng), dtype=np.uint8)
    items[4] = inst[:5].copy()           # a truncated instance as a whole tiny item
    items[6] = np.concatenate([inst, synth.rand_bytes(5, 0, 100 - len(inst))]) if len(inst) < 100 else items[6]
    return items


def check_batch(eng, items, res, flags=0):
    assert res.rc == 0
    assert res.n_items == len(items)
    n_rescan = 0
    n_total = 0
    for k, it in enumerate(items):
        single = eng.scan_host(it, flags=flags)
        assert single.rc == 0
        if res.state[k] == api.ITEM_RESCAN:
            n_rescan += 1
            assert len(res.matches(k)) == 0
            continue
        assert single.error == 0
        got = tuples(res.matches(k))
        # the records of one string are contiguous and ascending (YR_MATCHES list order); plain strings
        # come string-major, the lists of chain heads after them
        seen, last = set(), (None, -1)
        for s_, o, _, _ in got:
            if s_ != last[0]:
                assert s_ not in seen
                seen.add(s_)
            else:
                assert o > last[1]
            last = (s_, o)
        assert sorted(got) == canon(single.matches), "item %d" % k
        n_total += len(got)
    assert n_rescan == res.n_rescan
    return n_total, n_rescan


@pytest.mark.parametrize("kernel", [api.KERNEL_FILTER, api.KERNEL_WALK])
def test_packed_batch_equals_single_scans(setup, kernel):
    rs, image = setup
    eng = Engine(image, 0)
    eng.set_option(api.OPT_SCAN_KERNEL, kernel)
    items = make_items(rs)
    res = eng.scan_batch(items)
    n_total, n_rescan = check_batch(eng, items, res)
    assert n_total > 50
    # chained strings are resolved inside the batch, item by item
    assert n_rescan == 0
    assert res.stats["scan_kernel"] == kernel
    eng.close()


def test_batch_against_oracle(setup):
    rs, image = setup
    eng = Engine(image, 0)
    items = make_items(rs, seed=9)[:12]
    res = eng.scan_batch(items)
    assert res.rc == 0
    for k, it in enumerate(items):
        if res.state[k] == api.ITEM_RESCAN:
            continue
        want = yrh.oracle_scan(image, it)
        assert want.error == 0
        assert canon(res.matches(k)) == canon(want.matches)
    eng.close()


def test_one_allocation_batch(setup):
    """items laid out inside one pinned allocation keep their layout and are copied in slices"""
    rs, image = setup
    eng = Engine(image, 0)
    eng.set_option(api.OPT_H2D_CHUNK, 1 << 20)   # several slices
    items = make_items(rs, seed=4)
    total = sum((len(i) + 63) & ~63 for i in items) + 64
    pin = PinnedBuffer(total)
    arena = pin.array
    arena[:] = 0x41
    views, o = [], 5     # deliberately unaligned starts
    for it in items:
        arena[o : o + len(it)] = it
        views.append(arena[o : o + len(it)])
        o += len(it) + (len(it) * 7) % 13
    res = eng.scan_batch(views, one_allocation=True)
    n_total, _ = check_batch(eng, views, res)
    assert n_total > 50
    # same items, not declared as one allocation
    res2 = eng.scan_batch(views)
    assert res2.rc == 0 and (res2.state == res.state).all()
    assert [canon(res2.matches(k)) for k in range(len(views))] == [canon(res.matches(k)) for k in range(len(views))]
    eng.close()


def test_batch_without_chains_needs_no_rescan():
    rs = synth.ruleset(2, 1000)
    image = yrh.b200().compile(rs.source).flatten()
    eng = Engine(image, 0)
    items = [synth.corpus(rs, 65536 + 1000 * k, 4096, seed=k, plant_seed=50 + k) for k in range(40)]
    res = eng.scan_batch(items)
    n_total, n_rescan = check_batch(eng, items, res)
    assert n_rescan == 0 and n_total > 400
    # one launch sequence for the whole batch
    assert res.stats["kernel_launches"] < 40
    eng.close()


def test_fast_mode_and_empty_batches(setup):
    rs, image = setup
    eng = Engine(image, 0)
    items = make_items(rs)[:6]
    res = eng.scan_batch(items, flags=1)       # SCAN_FLAGS_FAST_MODE: order dependent, per-item path only
    assert res.rc == 0 and res.n_rescan == len(items) and (res.state == api.ITEM_RESCAN).all()
    res = eng.scan_batch([])
    assert res.rc == 0 and res.n_items == 0 and len(res.all) == 0
    res = eng.scan_batch([np.zeros(0, np.uint8), np.zeros(0, np.uint8)])
    assert res.rc == 0 and res.n_rescan == 0 and len(res.all) == 0
    eng.close()


def test_verifier_error_marks_only_its_item():
    """ERROR_TOO_MANY_RE_FIBERS raised inside one item must not leak into the others"""
    src = 'rule a { strings: $a = /ab(c|d|e|f|g|h|i|j|k|l|m){1,300}(c|d|e|f|g|h|i|j|k|l|m){1,300}x/ $b = "needle_in_item" condition: any of them }'
    image = yrh.b200().compile(src).flatten()
    eng = Engine(image, 0)
    bad = np.frombuffer(b"ab" + b"cdefghijklm" * 200, dtype=np.uint8).copy()
    single = eng.scan_host(bad)
    good = np.frombuffer(b"....needle_in_item....", dtype=np.uint8).copy()
    items = [good, bad, good.copy(), synth.rand_bytes(1, 0, 5000).copy()]
    res = eng.scan_batch(items)
    assert res.rc == 0
    if single.error != 0:
        assert res.state[1] == api.ITEM_RESCAN
    assert res.state[0] == api.ITEM_DONE and res.state[2] == api.ITEM_DONE and res.state[3] == api.ITEM_DONE
    assert len(res.matches(0)) == 1 and len(res.matches(2)) == 1 and len(res.matches(3)) == 0
    eng.close()


def test_dropin_batch_is_transparent(setup):
    """yr_b200_batch_scan + ordinary yr_rules_scan_mem per buffer == reference libyara per buffer:
    same rule verdicts, same YR_MATCH lists (base, offset, length, data hash, xor key)."""
    rs, image = setup
    rb = yrh.b200().compile(rs.source)
    rr = yrh.ref().compile(rs.source)
    items = make_items(rs, seed=6)
    want = [rr.scan(it).key() for it in items]
    h0 = yrh.batch_hits(yrh.b200())
    with rb.batch(items) as b:
        assert b.rc == 0
        assert b.n_rescans == 0
        got = [rb.scan(it).key() for it in items]
        # served from the parked results, not scanned again (empty items have no address to be found by)
        assert yrh.batch_hits(yrh.b200()) - h0 >= sum(1 for it in items if len(it) > 0)
        # a buffer that was not part of the batch, and a prefix of one that was, are scanned normally
        other = synth.corpus(rs, 50000, 4096, seed=77, plant_seed=78)
        assert rb.scan(other).key() == rr.scan(other).key()
        assert rb.scan(items[10][:5000]).key() == rr.scan(items[10][:5000]).key()
        # fast mode never uses parked results
        assert rb.scan(items[9], flags=1).key() == rr.scan(items[9], flags=1).key()
    assert got == want
    # after the batch is gone everything is scanned on demand again
    assert [rb.scan(it).key() for it in items[:12]] == want[:12]


def test_dropin_batch_multithreaded(setup):
    from concurrent.futures import ThreadPoolExecutor
    rs, image = setup
    rb = yrh.b200().compile(rs.source)
    items = make_items(rs, seed=8)
    want = [rb.scan(it).key() for it in items]
    with rb.batch(items) as b:
        assert b.rc == 0
        with ThreadPoolExecutor(4) as ex:
            got = list(ex.map(lambda it: rb.scan(it).key(), items))
    assert got == want


def test_chain_parts_in_neighbouring_items_do_not_connect():
    """a chained string whose head lies at the end of one item and whose tail lies at the start of the
    next must not match; the same bytes inside ONE item must"""
    src = 'rule a { strings: $a = { 11 22 33 44 [300-400] 55 66 77 88 [250-300] 99 AA BB CC } condition: $a }'
    image = yrh.b200().compile(src).flatten()
    eng = Engine(image, 0)
    head, mid, tail = bytes.fromhex("11223344"), bytes.fromhex("55667788"), bytes.fromhex("99aabbcc")
    whole = np.frombuffer(b"." * 50 + head + b"." * 350 + mid + b"." * 260 + tail + b"." * 20, dtype=np.uint8).copy()
    cut1 = 50 + 4 + 100                      # inside the first gap
    cut2 = 50 + 4 + 350 + 4 + 100            # inside the second gap
    items = [whole[:cut1].copy(), whole[cut1:].copy(), whole.copy(), whole[:cut2].copy(), whole[cut2:].copy(), whole.copy()]
    res = eng.scan_batch(items)
    assert res.rc == 0 and res.n_rescan == 0
    counts = [len(res.matches(k)) for k in range(len(items))]
    assert counts == [0, 0, 1, 0, 0, 1]
    for k in (2, 5):
        m = res.matches(k)[0]
        assert int(m["offset"]) == 50 and int(m["match_length"]) == 4 + 350 + 4 + 260 + 4
        single = eng.scan_host(items[k])
        assert canon(single.matches) == canon(res.matches(k))
    eng.close()


def test_file_sharding_over_ranks_equals_single_scans(setup):
    """shard.plan_files + shard.scan_files for 3 (simulated) ranks: every file scanned exactly once,
    records equal to a scan of the file alone"""
    from yara_b200 import shard
    rs, image = setup
    eng = Engine(image, 0)
    items = make_items(rs, seed=12)
    got = {}
    for ix in shard.plan_files([len(i) for i in items], 3):
        part = shard.scan_files(eng, items, ix)
        assert not (set(part) & set(got))
        got.update(part)
    assert sorted(got) == list(range(len(items)))
    for k, it in enumerate(items):
        single = eng.scan_host(it)
        assert got[k][0] == single.error
        assert canon(got[k][1]) == canon(single.matches)
    eng.close()
