"""ctypes bindings: `Engine` over libyrb200.so (the C ABI of include/yara_b200.h)
and `LibYara` over libyara_b200.so (the drop-in libyara: yr_compiler_*,
yr_rules_scan_mem ... unchanged public API, GPU block scan underneath)."""
import ctypes as C
import os

import numpy as np

YRB_MATCH_DTYPE = np.dtype(
    [("offset", "<i8"), ("string_idx", "<u4"), ("match_length", "<i4"), ("xor_key", "u1"), ("flags", "u1"),
     ("seed", "<u2"), ("chain_length", "<i4")]
)
RAW_DTYPE = np.dtype(
    [("key", "<u8"), ("offset", "<i8"), ("string_idx", "<u4"), ("length", "<i4"), ("misc", "<u4"), ("group", "<u4")]
)

KERNEL_AUTO, KERNEL_WALK, KERNEL_FILTER = 0, 1, 2
OPT_SCAN_KERNEL, OPT_H2D_CHUNK = 1, 2


def lib_dir():
    return os.path.dirname(os.path.abspath(__file__))


class ScanParams(C.Structure):
    _fields_ = [
        ("base", C.c_uint64),
        ("scan_flags", C.c_int32),
        ("too_many_action", C.c_int32),
        ("data_offset", C.c_uint64),
        ("block_size", C.c_uint64),
        ("own_begin", C.c_uint64),
        ("own_end", C.c_uint64),
        ("defer_resolve", C.c_int32),
        ("n_seeds", C.c_uint32),
        ("seeds", C.c_void_p),
        ("timeout_ns", C.c_uint64),
        ("n_skip_strings", C.c_uint32),
        ("reserved2", C.c_uint32),
        ("skip_strings", C.c_void_p),
    ]


class ScanStats(C.Structure):
    _fields_ = [
        ("bytes", C.c_uint64),
        ("n_candidates", C.c_uint64),
        ("n_raw", C.c_uint64),
        ("kernel_launches", C.c_uint32),
        ("scan_kernel", C.c_uint32),
        ("ms_h2d", C.c_float),
        ("ms_scan", C.c_float),
        ("ms_verify", C.c_float),
        ("ms_resolve", C.c_float),
        ("ms_total", C.c_float),
        ("ms_kernel", C.c_float),
        ("reserved", C.c_uint32),
    ]


class ScanResultC(C.Structure):
    _fields_ = [
        ("error", C.c_int32),
        ("error_string", C.c_uint32),
        ("n_matches", C.c_size_t),
        ("matches", C.c_void_p),
        ("n_unconfirmed", C.c_size_t),
        ("unconfirmed", C.c_void_p),
        ("n_too_many", C.c_uint32),
        ("too_many", C.POINTER(C.c_uint32)),
        ("seeds_consumed", C.c_uint32),
        ("reserved", C.c_uint32),
        ("stats", ScanStats),
    ]


class FilterInfo(C.Structure):
    _fields_ = [("usable", C.c_int32), ("variant", C.c_int32), ("n_windows", C.c_uint32), ("fill3", C.c_float),
                ("has_depth1", C.c_int32), ("has_depth2", C.c_int32), ("root_match", C.c_int32), ("reserved", C.c_int32)]


class BatchItem(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_uint64)]


class BatchResultC(C.Structure):
    _fields_ = [
        ("n_items", C.c_uint32),
        ("n_rescan", C.c_uint32),
        ("item_state", C.POINTER(C.c_uint8)),
        ("item_first", C.POINTER(C.c_uint64)),
        ("matches", C.c_void_p),
        ("stats", ScanStats),
    ]


ITEM_DONE, ITEM_RESCAN = 0, 1
BATCH_ONE_ALLOCATION = 1


class BatchResult:
    """yrb_batch_result copied out: state[k], and matches(k) = the item's yrb_match records."""

    def __init__(self, rc, r: BatchResultC):
        self.rc = rc
        self.n_items = r.n_items
        self.n_rescan = r.n_rescan
        self.stats = {f[0]: getattr(r.stats, f[0]) for f in ScanStats._fields_}
        n = r.n_items
        self.state = np.ctypeslib.as_array(r.item_state, shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        self.first = np.ctypeslib.as_array(r.item_first, shape=(n + 1,)).copy() if n else np.zeros(1, np.uint64)
        total = int(self.first[-1])
        if total and r.matches:
            self.all = np.frombuffer(C.string_at(r.matches, total * YRB_MATCH_DTYPE.itemsize), dtype=YRB_MATCH_DTYPE).copy()
        else:
            self.all = np.zeros(0, dtype=YRB_MATCH_DTYPE)

    def matches(self, k):
        return self.all[int(self.first[k]): int(self.first[k + 1])]


class ScanResult:
    def __init__(self, rc, r: ScanResultC, copy=True):
        self.rc = rc
        self.error = r.error
        self.error_string = r.error_string
        self.stats = {f[0]: getattr(r.stats, f[0]) for f in ScanStats._fields_}

        def arr(ptr, n):
            if not n or not ptr:
                return np.zeros(0, dtype=YRB_MATCH_DTYPE)
            return np.frombuffer(C.string_at(ptr, n * YRB_MATCH_DTYPE.itemsize), dtype=YRB_MATCH_DTYPE).copy()

        self.n_matches = r.n_matches
        self.matches = arr(r.matches, r.n_matches) if copy else None
        self.unconfirmed = arr(r.unconfirmed, r.n_unconfirmed) if copy else None
        self.too_many = [r.too_many[i] for i in range(r.n_too_many)]


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: code %d%s" % (what, rc, " (no CUDA device: the scan path has no CPU fallback)" if rc == 1001 else ""))


class _Lib:
    _lib = None

    @classmethod
    def get(cls):
        if cls._lib is None:
            path = os.path.join(lib_dir(), "libyrb200.so")
            if not os.path.exists(path):
                raise RuntimeError("libyrb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
            L = C.CDLL(path, mode=os.RTLD_GLOBAL | os.RTLD_NOW)
            L.yrb_version.restype = C.c_char_p
            L.yrb_rules_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
            L.yrb_rules_release.argtypes = [C.c_void_p]
            L.yrb_scanner_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
            L.yrb_scanner_destroy.argtypes = [C.c_void_p]
            L.yrb_scanner_set_option.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
            L.yrb_scan_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ScanParams), C.POINTER(ScanResultC)]
            L.yrb_scan_device.argtypes = L.yrb_scan_host.argtypes
            L.yrb_scan_batch.argtypes = [C.c_void_p, C.POINTER(BatchItem), C.c_uint32, C.c_int, C.c_int, C.POINTER(BatchResultC)]
            L.yrb_shard_raw.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
            L.yrb_resolve_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ScanParams), C.POINTER(ScanResultC)]
            L.yrb_gather_match_data.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p]
            L.yrb_dev_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
            L.yrb_dev_free.argtypes = [C.c_int, C.c_void_p]
            L.yrb_dev_fill_random.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
            L.yrb_dev_write.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            L.yrb_dev_read.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            L.yrb_dev_flush_l2.argtypes = [C.c_int]
            L.yrb_dev_scatter.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
            L.yrb_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
            L.yrb_host_free.argtypes = [C.c_void_p]
            L.yrb_multi_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
            L.yrb_multi_destroy.argtypes = [C.c_void_p]
            L.yrb_multi_device_count.argtypes = [C.c_void_p]
            L.yrb_multi_margin.argtypes = [C.c_void_p]
            L.yrb_multi_margin.restype = C.c_uint64
            L.yrb_multi_plan.argtypes = [C.c_void_p, C.c_uint64] + [C.POINTER(C.c_uint64)] * 4
            L.yrb_multi_scan_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ScanParams), C.POINTER(ScanResultC)]
            L.yrb_multi_scan_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)] + [C.POINTER(C.c_uint64)] * 4 + [
                C.c_uint64, C.POINTER(ScanParams), C.POINTER(ScanResultC)]
            L.yrb_filter_cache_stats.argtypes = [C.POINTER(C.c_uint)] * 3
            L.yrb_rules_filter_info.argtypes = [C.c_void_p, C.POINTER(FilterInfo)]
            L.yrb_shard_export_create.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            L.yrb_shard_export_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
            L.yrb_shard_export_publish.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
            L.yrb_shard_import_open.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
            L.yrb_shard_import_close.argtypes = [C.c_int, C.c_void_p]
            L.yrb_resolve_raw_device.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                                 C.POINTER(ScanParams), C.POINTER(ScanResultC)]
            cls._lib = L
        return cls._lib


class DeviceBuffer:
    """A block resident in HBM (yrb_dev_alloc)."""

    def __init__(self, device, size):
        self.L = _Lib.get()
        self.device, self.size = device, size
        p = C.c_void_p()
        _check(self.L.yrb_dev_alloc(device, size, C.byref(p)), "yrb_dev_alloc")
        self.ptr = p.value

    def fill_random(self, seed, stream_offset=0):
        _check(self.L.yrb_dev_fill_random(self.device, self.ptr, self.size, seed, stream_offset), "yrb_dev_fill_random")

    def write(self, offset, data):
        a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        _check(self.L.yrb_dev_write(self.device, self.ptr, offset, a.ctypes.data_as(C.c_void_p), a.size), "yrb_dev_write")

    def read(self, offset, n):
        out = np.empty(n, dtype=np.uint8)
        _check(self.L.yrb_dev_read(self.device, self.ptr, offset, out.ctypes.data_as(C.c_void_p), n), "yrb_dev_read")
        return out

    def read_into(self, offset, n, host_ptr):
        _check(self.L.yrb_dev_read(self.device, self.ptr, offset, C.c_void_p(host_ptr), n), "yrb_dev_read")

    def scatter(self, plants, local_start=0):
        """plants: [(block_offset, bytes)]; written at block_offset - local_start when inside the buffer."""
        pl = [(o - local_start, b) for o, b in plants if o - local_start >= 0 and o - local_start + len(b) <= self.size]
        if not pl:
            return 0
        offs = np.array([o for o, _ in pl], dtype=np.uint64)
        starts = np.zeros(len(pl) + 1, dtype=np.uint64)
        starts[1:] = np.cumsum([len(b) for _, b in pl])
        blob = np.frombuffer(b"".join(b for _, b in pl), dtype=np.uint8)
        _check(self.L.yrb_dev_scatter(self.device, self.ptr, self.size, len(pl), offs.ctypes.data_as(C.c_void_p),
                                      starts.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p)), "yrb_dev_scatter")
        return len(pl)

    def free(self):
        if self.ptr:
            self.L.yrb_dev_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory (cudaHostAlloc) exposed as a numpy array."""

    def __init__(self, size):
        self.L = _Lib.get()
        p = C.c_void_p()
        _check(self.L.yrb_host_alloc(size, C.byref(p)), "yrb_host_alloc")
        self.ptr, self.size = p.value, size
        self.array = np.ctypeslib.as_array((C.c_uint8 * size).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self.L.yrb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """Device-resident rules + one scanner (stream, work areas) on `device`."""

    def __init__(self, image: bytes, device=0):
        self.L = _Lib.get()
        self.device = device
        self.image = image
        self._img = np.frombuffer(image, dtype=np.uint8)
        r = C.c_void_p()
        _check(self.L.yrb_rules_create(self._img.ctypes.data_as(C.c_void_p), len(image), device, C.byref(r)), "yrb_rules_create")
        self.rules = r
        s = C.c_void_p()
        _check(self.L.yrb_scanner_create(self.rules, C.byref(s)), "yrb_scanner_create")
        self.scanner = s

    @staticmethod
    def device_count():
        return _Lib.get().yrb_device_count()

    def set_option(self, opt, value):
        _check(self.L.yrb_scanner_set_option(self.scanner, opt, value), "yrb_scanner_set_option")

    def filter_info(self):
        fi = FilterInfo()
        _check(self.L.yrb_rules_filter_info(self.rules, C.byref(fi)), "yrb_rules_filter_info")
        return {f[0]: getattr(fi, f[0]) for f in FilterInfo._fields_ if f[0] != "reserved"}

    def _params(self, base=0, flags=0, too_many_action=0, data_offset=0, block_size=0, own=None, defer=False, timeout_ns=0):
        p = ScanParams()
        p.base, p.scan_flags, p.too_many_action = base, flags, too_many_action
        p.timeout_ns = timeout_ns
        p.data_offset, p.block_size = data_offset, block_size
        if own is not None:
            p.own_begin, p.own_end = own
        p.defer_resolve = 1 if defer else 0
        return p

    def scan_host(self, data, copy=True, **kw):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_scan_host(self.scanner, a.ctypes.data_as(C.c_void_p), a.size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def scan_host_ptr(self, ptr, size, copy=True, **kw):
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_scan_host(self.scanner, C.c_void_p(ptr), size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def scan_batch(self, items, flags=0, one_allocation=False):
        """items: numpy uint8 arrays (kept alive by the caller) or (ptr, size) pairs."""
        arr = (BatchItem * max(1, len(items)))()
        for k, it in enumerate(items):
            if isinstance(it, tuple):
                arr[k].data, arr[k].size = it[0] or None, it[1]
            else:
                arr[k].data, arr[k].size = (it.ctypes.data if it.size else None), it.size
        r = BatchResultC()
        rc = self.L.yrb_scan_batch(self.scanner, arr, len(items), flags, BATCH_ONE_ALLOCATION if one_allocation else 0, C.byref(r))
        return BatchResult(rc, r)

    def scan_device(self, dbuf, size=None, copy=True, **kw):
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_scan_device(self.scanner, C.c_void_p(dbuf.ptr), dbuf.size if size is None else size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def shard_raw(self):
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(self.L.yrb_shard_raw(self.scanner, C.byref(ptr), C.byref(n)), "yrb_shard_raw")
        if n.value == 0:
            return np.zeros(0, dtype=RAW_DTYPE)
        return np.frombuffer(C.string_at(ptr, n.value * RAW_DTYPE.itemsize), dtype=RAW_DTYPE).copy()

    def resolve_raw(self, raw, **kw):
        raw = np.ascontiguousarray(raw)
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_resolve_raw(self.scanner, raw.ctypes.data_as(C.c_void_p), len(raw), C.byref(p), C.byref(r))
        return ScanResult(rc, r)

    # ---- shards scanned by different processes: records stay on the devices (CUDA IPC)
    def shard_export_create(self, capacity_records) -> bytes:
        """allocates this scanner's export buffer; returns its 64-byte IPC handle for the resolving process"""
        h = (C.c_uint8 * 64)()
        _check(self.L.yrb_shard_export_create(self.scanner, capacity_records, h, 64), "yrb_shard_export_create")
        return bytes(h)

    def shard_export_ptr(self) -> int:
        p = C.c_void_p()
        _check(self.L.yrb_shard_export_ptr(self.scanner, C.byref(p)), "yrb_shard_export_ptr")
        return p.value

    def shard_export_publish(self):
        """after a defer=True scan: (number of records now in the export buffer, any verifier error in the shard)"""
        n, e = C.c_size_t(), C.c_int()
        _check(self.L.yrb_shard_export_publish(self.scanner, C.byref(n), C.byref(e)), "yrb_shard_export_publish")
        return n.value, e.value

    def shard_import_open(self, handle: bytes) -> int:
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(self.L.yrb_shard_import_open(self.device, h, 64, C.byref(p)), "yrb_shard_import_open")
        return p.value

    def resolve_raw_device(self, ptrs, counts, copy=True, **kw):
        n = len(ptrs)
        pa = (C.c_void_p * max(1, n))(*ptrs)
        ca = (C.c_size_t * max(1, n))(*counts)
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_resolve_raw_device(self.scanner, n, pa, ca, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def gather_match_data(self, dbuf, n_matches, max_match_data=512):
        offs = np.zeros(n_matches + 1, dtype=np.uint64)
        out = np.empty(max(1, n_matches * max_match_data), dtype=np.uint8)
        _check(self.L.yrb_gather_match_data(self.scanner, C.c_void_p(dbuf.ptr), max_match_data, out.ctypes.data_as(C.c_void_p),
                                            out.size, offs.ctypes.data_as(C.c_void_p)), "yrb_gather_match_data")
        return out[: int(offs[-1])], offs

    def flush_l2(self):
        _check(self.L.yrb_dev_flush_l2(self.device), "yrb_dev_flush_l2")

    def close(self):
        if self.scanner:
            self.L.yrb_scanner_destroy(self.scanner)
            self.scanner = None
        if self.rules:
            self.L.yrb_rules_release(self.rules)
            self.rules = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiEngine:
    """ONE block over several GPUs from one process (yrb_multi_*): a rules replica and a
    scanner per device, byte-range shards with read margins, one global resolve on the first device."""

    def __init__(self, image: bytes, devices):
        self.L = _Lib.get()
        self.devices = list(devices)
        self._img = np.frombuffer(image, dtype=np.uint8)
        arr = (C.c_int * len(self.devices))(*self.devices)
        m = C.c_void_p()
        _check(self.L.yrb_multi_create(self._img.ctypes.data_as(C.c_void_p), len(image), arr, len(self.devices), C.byref(m)),
               "yrb_multi_create")
        self.m = m

    @property
    def margin(self):
        return int(self.L.yrb_multi_margin(self.m))

    def plan(self, block_size):
        """[(g0, g1, own0, own1)] per device"""
        n = len(self.devices)
        a = [(C.c_uint64 * n)() for _ in range(4)]
        _check(self.L.yrb_multi_plan(self.m, block_size, *a), "yrb_multi_plan")
        return [tuple(int(a[j][g]) for j in range(4)) for g in range(n)]

    def _params(self, base=0, flags=0, too_many_action=0, timeout_ns=0):
        p = ScanParams()
        p.base, p.scan_flags, p.too_many_action = base, flags, too_many_action
        p.timeout_ns = timeout_ns
        return p

    def scan_host(self, data, copy=True, **kw):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_multi_scan_host(self.m, a.ctypes.data_as(C.c_void_p), a.size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def scan_host_ptr(self, ptr, size, copy=True, **kw):
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_multi_scan_host(self.m, C.c_void_p(ptr), size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def scan_device(self, dbufs, plan, block_size, copy=True, **kw):
        """dbufs[g]: DeviceBuffer on device g holding block bytes [g0, g1) of plan[g]"""
        n = len(self.devices)
        ptrs = (C.c_void_p * n)(*[d.ptr for d in dbufs])
        cols = [(C.c_uint64 * n)(*[pl[j] for pl in plan]) for j in range(4)]
        p, r = self._params(**kw), ScanResultC()
        rc = self.L.yrb_multi_scan_device(self.m, ptrs, cols[0], cols[1], cols[2], cols[3], block_size, C.byref(p), C.byref(r))
        return ScanResult(rc, r, copy)

    def close(self):
        if self.m:
            self.L.yrb_multi_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def filter_cache_stats():
    a, b, c = C.c_uint(), C.c_uint(), C.c_uint()
    _Lib.get().yrb_filter_cache_stats(C.byref(a), C.byref(b), C.byref(c))
    return {"registry": a.value, "disk": b.value, "built": c.value}


# -------------------------------------------------------------------- libyara
YR_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)
CALLBACK_MSG_RULE_MATCHING = 1
CALLBACK_MSG_SCAN_FINISHED = 3


class LibYara:
    """The public libyara API of a libyara build (default: the GPU drop-in)."""

    def __init__(self, path=None):
        path = path or os.path.join(lib_dir(), "libyara_b200.so")
        if not os.path.exists(path):
            raise RuntimeError("%s is not built" % path)
        _Lib.get() if "b200" in os.path.basename(path) else None
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        L.yr_compiler_create.argtypes = [C.POINTER(C.c_void_p)]
        L.yr_compiler_destroy.argtypes = [C.c_void_p]
        L.yr_compiler_add_string.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.yr_compiler_get_rules.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.yr_rules_destroy.argtypes = [C.c_void_p]
        L.yr_rules_scan_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, YR_CALLBACK, C.c_void_p, C.c_int]
        L.yr_rules_save.argtypes = [C.c_void_p, C.c_char_p]
        L.yr_rules_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.yr_b200_rules_flatten.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.yr_b200_image_free.argtypes = [C.c_void_p]
        if hasattr(L, "yr_b200_set_device"):
            L.yr_b200_set_device.argtypes = [C.c_int]
        _check(L.yr_initialize(), "yr_initialize")

    def set_device(self, device):
        self.lib.yr_b200_set_device(device)

    def compile(self, source: str):
        comp = C.c_void_p()
        _check(self.lib.yr_compiler_create(C.byref(comp)), "yr_compiler_create")
        nerr = self.lib.yr_compiler_add_string(comp, source.encode("latin-1"), None)
        if nerr != 0:
            self.lib.yr_compiler_destroy(comp)
            raise ValueError("%d compile errors" % nerr)
        rules = C.c_void_p()
        rc = self.lib.yr_compiler_get_rules(comp, C.byref(rules))
        self.lib.yr_compiler_destroy(comp)
        _check(rc, "yr_compiler_get_rules")
        return rules

    def destroy(self, rules):
        self.lib.yr_rules_destroy(rules)

    def save(self, rules, path):
        """yr_rules_save (rules.c:433): the compiled-rules file `yarac` writes"""
        _check(self.lib.yr_rules_save(rules, path.encode()), "yr_rules_save")

    def load(self, path):
        """yr_rules_load (rules.c:372-431, arena.c:543)"""
        rules = C.c_void_p()
        _check(self.lib.yr_rules_load(path.encode(), C.byref(rules)), "yr_rules_load")
        return rules

    def flatten(self, rules) -> bytes:
        img, size = C.c_void_p(), C.c_size_t()
        _check(self.lib.yr_b200_rules_flatten(rules, C.byref(img), C.byref(size)), "yr_b200_rules_flatten")
        blob = C.string_at(img, size.value)
        self.lib.yr_b200_image_free(img)
        return blob

    def scan_mem(self, rules, ptr, size, flags=0, timeout=0):
        """yr_rules_scan_mem; returns (rc, number of matching rules)."""
        count = [0]

        def cb(ctx, msg, data, user):
            if msg == CALLBACK_MSG_RULE_MATCHING:
                count[0] += 1
            return 0

        fn = YR_CALLBACK(cb)
        rc = self.lib.yr_rules_scan_mem(rules, C.c_void_p(ptr), size, flags, fn, None, timeout)
        return rc, count[0]
