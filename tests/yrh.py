"""ctypes front-end of the parity harness (tests/harness/yrh.c) and of the CPU
oracle restatement (oracle/liboracle.so).  TEST INFRASTRUCTURE.

Two instances of the harness exist: `ref()` is linked to the reference libyara
built from /root/reference (oracle/_ref), `b200()` to the drop-in
yara_b200/libyara_b200.so whose block scan runs on the GPU.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MATCH_DTYPE = np.dtype(
    [
        ("base", "<i8"),
        ("offset", "<i8"),
        ("data_hash", "<u8"),
        ("string_idx", "<u4"),
        ("match_length", "<i4"),
        ("data_length", "<i4"),
        ("xor_key", "u1"),
        ("is_private", "u1"),
        ("pad", "<u2"),
    ]
)
assert MATCH_DTYPE.itemsize == 40

# include/yrb_image.h: yrb_match
YRB_MATCH_DTYPE = np.dtype(
    [
        ("offset", "<i8"),
        ("string_idx", "<u4"),
        ("match_length", "<i4"),
        ("xor_key", "u1"),
        ("flags", "u1"),
        ("seed", "<u2"),
        ("chain_length", "<i4"),
    ]
)
assert YRB_MATCH_DTYPE.itemsize == 24


def fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for b in data:
        h ^= b
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


class ScanResult:
    def __init__(self, rc, matches, rule_bits, too_many_calls):
        self.rc = rc
        self.matches = matches  # structured array, MATCH_DTYPE
        self.rule_bits = rule_bits
        self.too_many_calls = too_many_calls

    def key(self):
        """Everything parity compares, as plain python."""
        return (self.rc, self.matches.tobytes(), self.rule_bits.tobytes(), self.too_many_calls)


class Rules:
    def __init__(self, h, ptr):
        self.h = h
        self.ptr = ptr
        ns, nr, ts = C.c_uint32(), C.c_uint32(), C.c_uint32()
        h.lib.yrh_rules_info(ptr, C.byref(ns), C.byref(nr), C.byref(ts))
        self.num_strings, self.num_rules, self.tables_size = ns.value, nr.value, ts.value

    def __del__(self):
        try:
            if self.ptr:
                self.h.lib.yrh_rules_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass

    def flatten(self) -> bytes:
        img, size = C.c_void_p(), C.c_size_t()
        rc = self.h.lib.yrh_flatten(self.ptr, C.byref(img), C.byref(size))
        assert rc == 0, rc
        blob = C.string_at(img, size.value)
        self.h.lib.yrh_free_image(img)
        return blob

    def _finish(self, rc, out, n, bits, tm):
        assert n.value <= len(out), "harness output buffer too small"
        return ScanResult(rc, out[: n.value].copy(), bits, tm.value)

    def scan(self, data, flags=0, timeout=0, too_many_action=0, cap=1 << 16):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        bits = np.zeros(max(self.num_rules, 1), dtype=np.uint8)
        n, tm = C.c_size_t(), C.c_int()
        rc = self.h.lib.yrh_scan(
            self.ptr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), flags, timeout, too_many_action,
            out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n), bits.ctypes.data_as(C.c_void_p), C.byref(tm))
        if n.value > cap:
            return self.scan(data, flags, timeout, too_many_action, cap=n.value)
        return self._finish(rc, out, n, bits, tm)

    def scan_timeout_ns(self, data, timeout_ns):
        """yr_scanner_scan_mem with scanner->timeout set in nanoseconds; returns the scan's return code"""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        self.h.lib.yrh_scan_timeout_ns.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
        return self.h.lib.yrh_scan_timeout_ns(self.ptr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), timeout_ns)

    def scan_blocks(self, data, block_size, overlap, flags=0, too_many_action=0, cap=1 << 16):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        bits = np.zeros(max(self.num_rules, 1), dtype=np.uint8)
        n, tm = C.c_size_t(), C.c_int()
        rc = self.h.lib.yrh_scan_blocks(
            self.ptr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), C.c_size_t(block_size),
            C.c_size_t(overlap), flags, too_many_action, out.ctypes.data_as(C.c_void_p), C.c_size_t(cap),
            C.byref(n), bits.ctypes.data_as(C.c_void_p), C.byref(tm))
        if n.value > cap:
            return self.scan_blocks(data, block_size, overlap, flags, too_many_action, cap=n.value)
        return self._finish(rc, out, n, bits, tm)

    def scan_blocks_after_timed_out_scan(self, data1, slow_from, timeout_ns, data2, block_size, cap=1 << 16):
        """ONE scanner: a multi-block scan of data1 whose iterator turns slow at `slow_from` (the scan runs out of its
        time budget part-way), then a multi-block scan of data2 to the end.  Returns (rc of the first scan, result of
        the second)."""
        b1 = np.frombuffer(data1, dtype=np.uint8) if not isinstance(data1, np.ndarray) else data1
        b2 = np.frombuffer(data2, dtype=np.uint8) if not isinstance(data2, np.ndarray) else data2
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        bits = np.zeros(max(self.num_rules, 1), dtype=np.uint8)
        n, tm, rc1 = C.c_size_t(), C.c_int(), C.c_int()
        rc = self.h.lib.yrh_scan_blocks_after_timed_out_scan(
            self.ptr, b1.ctypes.data_as(C.c_void_p), C.c_size_t(b1.size), C.c_size_t(slow_from), C.c_uint64(timeout_ns),
            b2.ctypes.data_as(C.c_void_p), C.c_size_t(b2.size), C.c_size_t(block_size), out.ctypes.data_as(C.c_void_p),
            C.c_size_t(cap), C.byref(n), bits.ctypes.data_as(C.c_void_p), C.byref(rc1))
        assert n.value <= cap
        return rc1.value, self._finish(rc, out, n, bits, tm)

    def scanner_sequence(self, scans, cap=1 << 18):
        """ONE YR_SCANNER over `scans` = [(data, mode, flags)]: mode 0 = yr_scanner_scan_mem, else yr_scanner_scan_mem_blocks
        with blocks of `mode` bytes.  Returns [(rc, matches, rule_bits)] per scan."""
        n = len(scans)
        bufs = [np.ascontiguousarray(np.frombuffer(d, dtype=np.uint8) if not isinstance(d, np.ndarray) else d) for d, _, _ in scans]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        modes = (C.c_int * n)(*[m for _, m, _ in scans])
        flags = (C.c_int * n)(*[f for _, _, f in scans])
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        per = (C.c_size_t * n)()
        rcs = (C.c_int * n)()
        bits = np.zeros(n * max(self.num_rules, 1), dtype=np.uint8)
        rc = self.h.lib.yrh_scanner_sequence(self.ptr, n, ptrs, sizes, modes, flags, out.ctypes.data_as(C.c_void_p), C.c_size_t(cap),
                                             per, bits.ctypes.data_as(C.c_void_p), rcs)
        assert rc == 0 and sum(per) <= cap
        res, at = [], 0
        for k in range(n):
            res.append((rcs[k], out[at:at + per[k]].copy(), bits[k * self.num_rules:(k + 1) * self.num_rules].copy()))
            at += per[k]
        return res

    def scan_regions(self, data, regions, flags=0, cap=1 << 18):
        """multi-block scan over regions = [(offset in data, size, base)], in that order"""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        n = len(regions)
        offs = (C.c_size_t * n)(*[r[0] for r in regions])
        sizes = (C.c_size_t * n)(*[r[1] for r in regions])
        bases = (C.c_uint64 * n)(*[r[2] for r in regions])
        out = np.zeros(cap, dtype=MATCH_DTYPE)
        bits = np.zeros(max(self.num_rules, 1), dtype=np.uint8)
        cnt, tm = C.c_size_t(), C.c_int()
        rc = self.h.lib.yrh_scan_regions(self.ptr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(n), offs, sizes, bases, flags,
                                         out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(cnt), bits.ctypes.data_as(C.c_void_p))
        assert cnt.value <= cap
        return self._finish(rc, out, cnt, bits, tm)

    def batch(self, arrays, flags=0, one_allocation=False):
        """GPU drop-in only: scan `arrays` as one batch; scans of the same arrays while the returned
        object is open take their matches from it (yara_b200/host/yr_b200.h)."""
        return Batch(self, arrays, flags, one_allocation)

    def scan_mt(self, data, nthreads, flags=0):
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        rc = C.c_int()
        secs = self.h.lib.yrh_scan_mt(
            self.ptr, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), nthreads, flags, C.byref(rc))
        return secs, rc.value


class Batch:
    def __init__(self, rules, arrays, flags, one_allocation):
        L = rules.h.lib
        L.yrh_batch_begin.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_uint32, C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        L.yrh_batch_end.argtypes = [C.c_void_p]
        n = len(arrays)
        self.arrays = arrays          # keep the buffers alive
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in arrays])
        sizes = (C.c_size_t * max(n, 1))(*[a.size for a in arrays])
        h, nres = C.c_void_p(), C.c_uint32()
        self.rc = L.yrh_batch_begin(rules.ptr, ptrs, sizes, n, flags, 1 if one_allocation else 0, C.byref(h), C.byref(nres))
        self.lib, self.handle, self.n_rescans = L, h, nres.value

    def close(self):
        if self.handle:
            self.lib.yrh_batch_end(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Harness:
    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L = self.lib
        L.yrh_compile.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p]
        L.yrh_rules_destroy.argtypes = [C.c_void_p]
        L.yrh_rules_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 3
        L.yrh_flatten.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.yrh_free_image.argtypes = [C.c_void_p]
        L.yrh_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_int)]
        L.yrh_scan_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_int)]
        L.yrh_scan_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.yrh_scan_mt.restype = C.c_double
        L.yrh_set_max_match_data.argtypes = [C.c_uint32]
        rc = L.yrh_initialize()
        assert rc == 0, rc

    def compile(self, text: str) -> Rules:
        ptr = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.yrh_compile(text.encode("latin-1"), C.byref(ptr), err)
        if rc != 0:
            raise ValueError("compile failed rc=%d %s" % (rc, err.value.decode("latin-1", "replace")))
        return Rules(self, ptr)


_cache = {}


def ref_path():
    return os.path.join(ROOT, "oracle", "_ref", "libyrh_ref.so")


def b200_path():
    return os.path.join(ROOT, "yara_b200", "libyrh_b200.so")


def ref() -> Harness:
    if "ref" not in _cache:
        _cache["ref"] = Harness(ref_path())
    return _cache["ref"]


def b200cpu_path():
    return os.path.join(ROOT, "tests", "harness", "libyrh_b200cpu.so")


def b200cpu() -> Harness:
    """TEST stand-in: the drop-in's glue linked against tests/harness/fake_engine.cpp, which replays the
    product's verify / resolve device code on the CPU -- drives yr_b200_glue.c without a GPU"""
    if "b200cpu" not in _cache:
        _cache["b200cpu"] = Harness(b200cpu_path())
    return _cache["b200cpu"]


def refpf_path():
    return os.path.join(ROOT, "oracle", "_ref", "libyrh_refpf.so")


def refpf() -> Harness:
    """reference + the condition-VM pre-filter hooks only (CPU); the filter obeys $YR_B200_RULE_PREFILTER,
    read once at the first scan"""
    if "refpf" not in _cache:
        h = Harness(refpf_path())
        h.lib.yrh_prefilter_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        h.lib.yrh_prefilter_classes.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        _cache["refpf"] = h
    return _cache["refpf"]


def batch_hits(h):
    """scans the drop-in served from parked file-batch results so far"""
    h.lib.yrh_batch_hits.restype = C.c_uint64
    return h.lib.yrh_batch_hits()


def prefilter_stats(h):
    e, s = C.c_uint64(), C.c_uint64()
    h.lib.yrh_prefilter_stats(C.byref(e), C.byref(s))
    return e.value, s.value


def prefilter_classes(h, rules):
    c = (C.c_uint32 * 5)()
    h.lib.yrh_prefilter_classes(rules.ptr, c)
    return {"impure": c[0], "unknown": c[1], "skippable": c[2], "true_without_matches": c[3], "walked": c[4]}


def b200() -> Harness:
    """the GPU drop-in.  With YRH_B200_IMPL=cpu in the environment (a development aid: lets the harness-level
    GPU tests be dry-run without a GPU) the stand-in build of tests/harness/ is returned instead."""
    if os.environ.get("YRH_B200_IMPL") == "cpu":
        return b200cpu()
    if "b200" not in _cache:
        _cache["b200"] = Harness(b200_path())
    return _cache["b200"]


# ------------------------------------------------------------ oracle restatement
class _YroResult(C.Structure):
    _fields_ = [
        ("error", C.c_int),
        ("error_string", C.c_uint32),
        ("n_matches", C.c_size_t),
        ("matches", C.c_void_p),
        ("n_unconfirmed", C.c_size_t),
        ("unconfirmed", C.c_void_p),
        ("n_too_many", C.c_uint32),
        ("too_many", C.POINTER(C.c_uint32)),
        ("n_candidates", C.c_uint64),
        ("max_fibers", C.c_uint64),
        ("max_positions", C.c_uint64),
    ]


class OracleResult:
    def __init__(self, error, matches, unconfirmed, too_many, n_candidates, max_fibers, max_positions):
        self.error = error
        self.matches = matches  # YRB_MATCH_DTYPE
        self.unconfirmed = unconfirmed
        self.too_many = too_many
        self.n_candidates = n_candidates
        self.max_fibers = max_fibers
        self.max_positions = max_positions


def _oracle_lib():
    if "oracle" not in _cache:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        lib.yro_scan_block.restype = C.POINTER(_YroResult)
        lib.yro_scan_block.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_int]
        lib.yro_result_free.argtypes = [C.POINTER(_YroResult)]
        lib.yro_ac_events.restype = C.c_size_t
        lib.yro_ac_events.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        _cache["oracle"] = lib
    return _cache["oracle"]


def oracle_scan(image: bytes, data, base=0, flags=0, too_many_action=0) -> OracleResult:
    lib = _oracle_lib()
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    img = np.frombuffer(image, dtype=np.uint8)
    rp = lib.yro_scan_block(img.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size),
                            C.c_uint64(base), flags, too_many_action)
    assert rp, "bad image"
    r = rp.contents

    def arr(ptr, n):
        if n == 0:
            return np.zeros(0, dtype=YRB_MATCH_DTYPE)
        return np.frombuffer(C.string_at(ptr, n * YRB_MATCH_DTYPE.itemsize), dtype=YRB_MATCH_DTYPE).copy()

    out = OracleResult(r.error, arr(r.matches, r.n_matches), arr(r.unconfirmed, r.n_unconfirmed),
                       [r.too_many[i] for i in range(r.n_too_many)], r.n_candidates, r.max_fibers, r.max_positions)
    lib.yro_result_free(rp)
    return out


def oracle_ac_events(image: bytes, data, cap=1 << 20):
    lib = _oracle_lib()
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    img = np.frombuffer(image, dtype=np.uint8)
    pos = np.zeros(cap, dtype=np.uint64)
    acm = np.zeros(cap, dtype=np.uint32)
    n = lib.yro_ac_events(img.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size),
                          pos.ctypes.data_as(C.c_void_p), acm.ctypes.data_as(C.c_void_p), C.c_size_t(cap))
    if n > cap:
        return oracle_ac_events(image, data, cap=n)
    return pos[:n].copy(), acm[:n].copy()


def matches_from_flat(flat, data, max_match_data=512, base=0, private_mask=None):
    """Turn engine/oracle yrb_match records into the harness's MATCH_DTYPE rows
    (adds base, data_length, data hash) so both sides compare byte for byte."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    out = np.zeros(len(flat), dtype=MATCH_DTYPE)
    out["base"] = base
    out["offset"] = flat["offset"]
    out["string_idx"] = flat["string_idx"]
    out["match_length"] = flat["match_length"]
    out["xor_key"] = flat["xor_key"]
    dl = np.minimum(flat["match_length"], max_match_data).astype(np.int32)
    out["data_length"] = dl
    if private_mask is not None:
        out["is_private"] = private_mask[flat["string_idx"]]
    mv = memoryview(buf)
    for i in range(len(flat)):
        o = int(flat["offset"][i])
        out["data_hash"][i] = fnv1a(bytes(mv[o : o + max(int(dl[i]), 0)]))
    return out
