"""CPU replay of the PRODUCT's verifier code against the oracle.

tests/harness/vm_host.cpp compiles yara_b200/csrc/verify.cu and emit.cuh -- candidate emission, the
literal compare family, the fast hex VM, the fiber VM and their work-area tiers, error reporting --
for the CPU (tests/harness/cuda_host_shim.h) and runs the kernels as ordinary functions, one thread
of a 1x1 grid.  What comes out (match-callback records in discovery-key order, the earliest verifier
error) must equal the oracle's callback trace of the same block: the same device code the GPU runs,
checked without a GPU.  (The scan kernels and the resolve stage use warp collectives / CUB and stay
with the -m gpu tests.)"""
import ctypes as C
import os

import numpy as np
import pytest

import fuzzgen
import yrh
from cases import CASES
from yara_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "harness", "libvmhost.so")
HAVE = os.path.exists(LIB) and os.path.exists(yrh.ref_path())
pytestmark = pytest.mark.skipif(not HAVE, reason="tests/harness/libvmhost.so or oracle/_ref not built")

RAW = np.dtype([("key", "<u8"), ("offset", "<i8"), ("string_idx", "<u4"), ("length", "<i4"), ("xor_key", "<u4"), ("group", "<u4")])
CB = np.dtype([("offset", "<i8"), ("string_idx", "<u4"), ("length", "<i4"), ("xor_key", "<u4"), ("pad", "<u4")])


class Stats(C.Structure):
    _fields_ = [("n_cand", C.c_uint64 * 3), ("n_raw", C.c_uint64), ("retry_fast", C.c_uint32), ("retry_re", C.c_uint32),
                ("retry_fast3", C.c_uint32), ("retry_fast4", C.c_uint32), ("error", C.c_int32), ("error_string", C.c_uint32),
                ("error_key", C.c_uint64)]


_lib = None


def vmlib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.vmh_run.restype = C.c_longlong
        _lib.vmh_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Stats)]
    return _lib


def vm_run(image, data, cap=1 << 16):
    img = np.frombuffer(image, dtype=np.uint8)
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    out = np.zeros(cap, dtype=RAW)
    st = Stats()
    n = vmlib().vmh_run(img.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), buf.size, 0, 0,
                        out.ctypes.data_as(C.c_void_p), cap, C.byref(st))
    assert n >= 0
    if n > cap:
        return vm_run(image, data, cap=int(n))
    return out[:n], st


def oracle_trace(image, data, cap=1 << 16):
    lib = yrh._oracle_lib()
    lib.yro_trace_begin.argtypes = [C.c_void_p, C.c_size_t]
    lib.yro_trace_end.restype = C.c_size_t
    buf = np.zeros(cap, dtype=CB)
    lib.yro_trace_begin(buf.ctypes.data_as(C.c_void_p), cap)
    o = yrh.oracle_scan(image, data)
    n = lib.yro_trace_end()
    if n > cap:
        return oracle_trace(image, data, cap=int(n))
    return buf[:n], o


def check(image, data):
    want, o = oracle_trace(image, data)
    got, st = vm_run(image, data)
    if o.error == 0:
        assert st.error == 0
    else:
        # the reference stops at its first failing candidate; the product verifies every candidate and keeps
        # the earliest failure: same code, and everything discovered before it is the same
        assert st.error == o.error
        got = got[got["key"] < st.error_key]
    n = len(got)
    if o.error != 0:
        n = min(len(got), len(want))   # callbacks of the failing candidate itself may precede its error
    else:
        assert len(got) == len(want), (len(got), len(want))
    for f in ("string_idx", "offset", "length", "xor_key"):
        assert np.array_equal(got[f][:n].astype(np.int64), want[f][:n].astype(np.int64)), f
    b = want[:n]
    return len(b), st


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases(idx):
    rule, data = CASES[idx]
    image = yrh.ref().compile(rule).flatten()
    check(image, data)


FUZZ_SEEDS = [s for s in range(24) if s != 19]   # 19: half a minute of fiber VM on the CPU; scripts/vm_host_sweep.py runs it


@pytest.mark.parametrize("seed", FUZZ_SEEDS)
def test_fuzz_rulesets(seed):
    R = yrh.ref()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 60)
    image = R.compile(text).flatten()
    data = fuzzgen.gen_data(seed, insts, (1 << 15) + 13 * seed)
    n, st = check(image, data)
    assert n > 0


@pytest.mark.parametrize("cfg,n", [(2, 500), (3, 2000), (4, 2000), (5, 800)])
def test_synthetic_rulesets(cfg, n):
    rs = synth.ruleset(cfg, n)
    image = yrh.ref().compile(rs.source).flatten()
    data = synth.corpus(rs, 1 << 19, 512)
    nrec, st = check(image, data)
    assert nrec > 500


def test_tiers_are_exercised():
    """the known-answer cases include position lists / fiber counts that outgrow tier 1, tier 2 and tier 3:
    the replay takes the same escalation path as the GPU"""
    R = yrh.ref()
    seen = {"retry_fast": 0, "retry_re": 0, "retry_fast3": 0}
    for rule, data in CASES:
        _, st = vm_run(R.compile(rule).flatten(), data)
        for k in seen:
            seen[k] += getattr(st, k)
    assert seen["retry_fast"] > 0 and seen["retry_re"] > 0 and seen["retry_fast3"] > 0, seen


# ------------------------------------------------------------------ verification + resolve
class Result(C.Structure):
    _fields_ = [("error", C.c_int32), ("error_string", C.c_uint32), ("n_matches", C.c_uint64), ("n_unconfirmed", C.c_uint64),
                ("n_too_many", C.c_uint32), ("over_cap", C.c_uint32), ("first_cut_key", C.c_uint64), ("error_key", C.c_uint64),
                ("n_cuts", C.c_uint32), ("pad", C.c_uint32)]


def vm_scan(image, data, flags=0, cap=1 << 16):
    """the product's verify + resolve device code on the CPU: what yrb_scan_host returns for one block"""
    from yara_b200.api import YRB_MATCH_DTYPE
    lib = vmlib()
    lib.vmh_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                             C.c_size_t, C.POINTER(Result), C.POINTER(Stats)]
    img = np.frombuffer(image, dtype=np.uint8)
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    m, u = np.zeros(cap, dtype=YRB_MATCH_DTYPE), np.zeros(cap, dtype=YRB_MATCH_DTYPE)
    res, st = Result(), Stats()
    rc = lib.vmh_scan(img.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), buf.size, 0, flags,
                      m.ctypes.data_as(C.c_void_p), cap, u.ctypes.data_as(C.c_void_p), cap, C.byref(res), C.byref(st))
    assert rc == 0
    if res.n_matches > cap or res.n_unconfirmed > cap:
        return vm_scan(image, data, flags, cap=int(max(res.n_matches, res.n_unconfirmed)))
    return m[: res.n_matches], u[: res.n_unconfirmed], res


def canon(a, fields):
    """rows of the named fields as int64, sorted lexicographically (a list of tuples, cheaply)"""
    if len(a) == 0:
        return []
    t = np.stack([np.asarray(a[f]).astype(np.int64) for f in fields], axis=1)
    t = t[np.lexsort(t.T[::-1])]
    return t.tolist()


def check_scan(image, data, flags):
    o = yrh.oracle_scan(image, data, flags=flags)
    m, u, res = vm_scan(image, data, flags)
    assert not res.over_cap
    assert res.n_too_many == len(o.too_many)
    assert res.error == o.error
    if o.error != 0:
        return 0
    # the records of one string are contiguous and already in YR_MATCHES list order (ascending offsets)
    if len(m) > 1:
        sid, off = m["string_idx"].astype(np.int64), m["offset"].astype(np.int64)
        same = sid[1:] == sid[:-1]
        assert (off[1:][same] > off[:-1][same]).all()
        starts = np.concatenate([[True], ~same])
        assert len(np.unique(sid[starts])) == int(starts.sum())
    assert canon(m, ("string_idx", "offset", "match_length", "xor_key")) == canon(o.matches, ("string_idx", "offset", "match_length", "xor_key"))
    assert canon(u, ("string_idx", "offset", "match_length", "chain_length")) == canon(o.unconfirmed, ("string_idx", "offset", "match_length", "chain_length"))
    return len(m)


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_scan_known_answer_cases(idx):
    rule, data = CASES[idx]
    image = yrh.ref().compile(rule).flatten()
    for flags in (0, 1):
        check_scan(image, data, flags)


@pytest.mark.parametrize("seed", FUZZ_SEEDS)
def test_scan_fuzz_rulesets(seed):
    R = yrh.ref()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 60)
    image = R.compile(text).flatten()
    data = fuzzgen.gen_data(seed, insts, (1 << 15) + 13 * seed)
    for flags in (0, 1):
        check_scan(image, data, flags)


@pytest.mark.parametrize("cfg,n", [(2, 500), (3, 2000), (4, 2000), (5, 800)])
def test_scan_synthetic_rulesets(cfg, n):
    rs = synth.ruleset(cfg, n)
    image = yrh.ref().compile(rs.source).flatten()
    data = synth.corpus(rs, 1 << 19, 512)
    assert check_scan(image, data, 0) > 500
    check_scan(image, data, 1)


# ------------------------------------------------------------------ batch mode (yrb_scan_batch) on the CPU
def vm_scan_batch(image, items, rng):
    """the device side of yrb_scan_batch: one walk / verify / resolve over the concatenation; returns the
    per-item records (item-relative offsets) and the items flagged for the per-item path"""
    from yara_b200.api import YRB_MATCH_DTYPE
    lib = vmlib()
    lib.vmh_scan_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p, C.POINTER(Result), C.POINTER(Stats)]
    starts, parts, cur = [], [], 0
    for it in items:
        starts.append(cur)
        parts.append(it)
        cur += len(it)
        pad = (-cur) % 16 if rng.random() < 0.7 else rng.choice([0, 1, 5])     # the engine packs at 16-byte boundaries
        if pad:
            src = items[rng.randrange(len(items))]
            parts.append(np.resize(src, pad).astype(np.uint8) if len(src) else np.zeros(pad, np.uint8))
            cur += pad
    concat = np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(1, np.uint8)
    st_a = np.array(starts, dtype=np.int64)
    ln_a = np.array([len(i) for i in items], dtype=np.int64)
    dirty = np.zeros(max(1, len(items)), dtype=np.uint8)
    img = np.frombuffer(image, dtype=np.uint8)
    cap = 1 << 16
    while True:
        m = np.zeros(cap, dtype=YRB_MATCH_DTYPE)
        res, st = Result(), Stats()
        rc = lib.vmh_scan_batch(img.ctypes.data_as(C.c_void_p), concat.ctypes.data_as(C.c_void_p), cur, len(items),
                                st_a.ctypes.data_as(C.c_void_p), ln_a.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), cap,
                                dirty.ctypes.data_as(C.c_void_p), C.byref(res), C.byref(st))
        assert rc == 0
        if res.n_matches <= cap:
            break
        cap = int(res.n_matches)
        dirty[:] = 0
    m = m[: res.n_matches]
    owner = np.searchsorted(st_a, m["offset"], side="right") - 1
    per_item = []
    for k in range(len(items)):
        sel = m[owner == k].copy()
        sel["offset"] -= st_a[k]
        per_item.append(sel)
    return per_item, dirty, res


def check_batch(image, items, rng):
    per_item, dirty, res = vm_scan_batch(image, items, rng)
    if res.over_cap:
        pytest.skip("1M cap")
    n = 0
    for k, it in enumerate(items):
        o = yrh.oracle_scan(image, it)
        if dirty[k]:
            assert o.error != 0 or o.too_many, "item %d flagged without a reason" % k
            continue
        assert o.error == 0, "item %d: the oracle fails (%d) but the batch did not flag it" % (k, o.error)
        f = ("string_idx", "offset", "match_length", "xor_key")
        assert canon(per_item[k], f) == canon(o.matches, f), "item %d of %d" % (k, len(items))
        n += len(o.matches)
    return n


def ragged(data, rng, n):
    data = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    cuts = sorted(set([0, len(data)] + [rng.randrange(len(data) + 1) for _ in range(n)]))
    items = [np.ascontiguousarray(data[a:b]) for a, b in zip(cuts, cuts[1:])]
    items.insert(rng.randrange(len(items) + 1), np.zeros(0, np.uint8))
    items.append(np.ascontiguousarray(data))          # and the whole buffer as one more item
    return items


@pytest.mark.parametrize("seed", range(16))
def test_batch_fuzz_rulesets(seed):
    import random
    R = yrh.ref()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 50)
    image = R.compile(text).flatten()
    if np.frombuffer(image, dtype=np.uint32, count=1, offset=int(np.frombuffer(image[56:64], dtype=np.uint64)[0]))[0] != 0:
        pytest.skip("root-state matches: batches are refused")
    data = fuzzgen.gen_data(seed, insts, 1 << 14)
    rng = random.Random(seed)
    check_batch(image, ragged(data, rng, 12), rng)


@pytest.mark.parametrize("cfg,n", [(3, 1500), (4, 2000), (5, 600)])
def test_batch_synthetic_rulesets(cfg, n):
    import random
    rs = synth.ruleset(cfg, n)
    image = yrh.ref().compile(rs.source).flatten()
    data = synth.corpus(rs, 1 << 18, 512)
    rng = random.Random(cfg)
    assert check_batch(image, ragged(data, rng, 30), rng) > 300


def test_batch_chain_parts_in_neighbouring_items():
    import random
    src = "rule a { strings: $a = { 11 22 33 44 [300-400] 55 66 77 88 [250-300] 99 AA BB CC } condition: $a }"
    image = yrh.ref().compile(src).flatten()
    head, mid, tail = bytes.fromhex("11223344"), bytes.fromhex("55667788"), bytes.fromhex("99aabbcc")
    whole = np.frombuffer(b"." * 50 + head + b"." * 350 + mid + b"." * 260 + tail + b"." * 20, dtype=np.uint8).copy()
    c1, c2 = 50 + 4 + 100, 50 + 4 + 350 + 4 + 100
    items = [whole[:c1].copy(), whole[c1:].copy(), whole.copy(), whole[:c2].copy(), whole[c2:].copy(), whole.copy()]
    per_item, dirty, res = vm_scan_batch(image, items, random.Random(1))
    assert [len(x) for x in per_item] == [0, 0, 1, 0, 0, 1] and not dirty.any()
    check_batch(image, items, random.Random(2))


def test_scan_one_million_match_cap():
    """YR_MAX_STRING_MATCHES (scan.c:303-307): the cut key and the second resolve pass, replayed on the CPU"""
    R = yrh.ref()
    image = R.compile('rule t { strings: $a = "aa" $b = "ab" condition: any of them }').flatten()
    data = np.frombuffer(b"a" * 1000100 + b"ab" * 10, dtype=np.uint8)
    o = yrh.oracle_scan(image, data)
    assert len(o.too_many) == 1
    m, u, res = vm_scan(image, data, 0, cap=1 << 21)
    assert res.n_too_many == 1 and res.n_cuts == 1 and not res.over_cap
    f = ("string_idx", "offset", "match_length", "xor_key")
    assert len(m) == len(o.matches) and canon(m, f) == canon(o.matches, f)


# ------------------------------------------------------------------ long YR_AC_MATCH chains (DevRules::key_seq_bits)
@pytest.mark.parametrize("n,seq_bits", [(300, 17), (600, 16), (1200, 15)])
def test_long_match_chains(n, seq_bits):
    """Hundreds of rules sharing one string put hundreds of entries on one state's chain.  The chain position's
    share of the discovery key grows with the ruleset's longest chain (engine.cuh yrb_key_seq_bits); every candidate
    keeps a key of its own, in emission order (checked inside the replay: vmh_run fails otherwise), the callback
    trace in key order is the oracle's, and so are the final lists in both scan modes."""
    import long_chains
    image = yrh.ref().compile(long_chains.ruleset(n)).flatten()
    img = np.frombuffer(image, dtype=np.uint8)
    lib = vmlib()
    lib.vmh_longest_match_chain.restype = C.c_uint32
    lib.vmh_key_seq_bits.restype = C.c_uint32
    longest = lib.vmh_longest_match_chain(img.ctypes.data_as(C.c_void_p))
    assert longest >= n // 2                              # half of the kinds pick the same atom: one state's chain holds them all
    assert lib.vmh_key_seq_bits(img.ctypes.data_as(C.c_void_p)) == seq_bits
    assert (1 << (25 - seq_bits)) - 1 >= longest
    data = long_chains.corpus()
    cnt, st = check(image, data)
    assert cnt > 20 * n
    assert check_scan(image, data, 0) > 10 * n
    check_scan(image, data, 1)


def test_a_chain_field_too_narrow_for_the_ruleset_is_caught(monkeypatch):
    """The same replay with the chain position squeezed into the 8 bits that suffice for chains of up to 255
    entries: candidates further down the chain collide with (or overtake) their neighbours' keys, and the replay's
    own key check refuses to go on."""
    import long_chains
    image = yrh.ref().compile(long_chains.ruleset(1200)).flatten()
    img = np.frombuffer(image, dtype=np.uint8)
    data = long_chains.corpus()
    out = np.zeros(1 << 16, dtype=RAW)
    st = Stats()
    args = (img.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), data.size, 0, 0, out.ctypes.data_as(C.c_void_p), 1 << 16, C.byref(st))
    assert vmlib().vmh_run(*args) > 0
    monkeypatch.setenv("VMH_FORCE_KEY_SEQ_BITS", "17")
    assert vmlib().vmh_run(*args) == -3
