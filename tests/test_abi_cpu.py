"""CPU tests: the C-ABI library loads without a GPU, exports every symbol
include/yara_b200.h declares, and refuses to compute (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yara_b200", "libyrb200.so")
DROPIN = os.path.join(ROOT, "yara_b200", "libyara_b200.so")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "yara_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yrb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_dropin_exports_libyara_api_and_hooks():
    lib = ctypes.CDLL(DROPIN)
    for s in ("yr_initialize", "yr_compiler_create", "yr_compiler_add_string", "yr_compiler_get_rules", "yr_rules_scan_mem",
              "yr_scanner_create", "yr_scanner_scan_mem", "yr_scanner_scan_mem_blocks", "yr_rules_destroy",
              "yr_b200_scan_mem_block", "yr_b200_scanner_release", "yr_b200_rules_release", "yr_b200_rules_flatten"):
        assert hasattr(lib, s), s
    # the reference's CPU verification path is deliberately not linked into the drop-in
    assert not hasattr(lib, "yr_scan_verify_match")


def test_dropin_exports_the_whole_public_api_of_the_reference_build():
    """drop-in means drop-in: every yr_* symbol the reference library (built from the same sources by
    oracle/Makefile) exports is exported by libyara_b200.so too, except the CPU verification path"""
    import subprocess
    ref = os.path.join(ROOT, "oracle", "_ref", "libyara_ref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")

    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return set(l.split()[-1] for l in out.splitlines() if l.split() and l.split()[-1].startswith("yr_"))

    a, b = exported(ref), exported(DROPIN)
    missing = sorted(a - b)
    # scan.c is not linked (no CPU fallback): its three entry points are the only ones allowed to be absent
    assert set(missing) <= {"yr_scan_verify_match", "yr_scan_base64_callback_dummy"}, missing
    assert "yr_scan_verify_match" in missing
    assert len(a) > 150


def test_no_cpu_fallback():
    lib = ctypes.CDLL(LIB)
    if lib.yrb_device_count() > 0:
        pytest.skip("a GPU is present")
    img = np.fromfile(os.path.join(ROOT, "tests", "golden", "synth_cfg2.img"), dtype=np.uint8)
    out = ctypes.c_void_p()
    rc = lib.yrb_rules_create(img.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(img.size), 0, ctypes.byref(out))
    assert rc == 1001  # YRB_ERROR_NO_DEVICE

    import yrh
    r = yrh.b200().compile('rule t { strings: $a = "abc" condition: $a }')
    res = r.scan(b"xxabcxx")
    assert res.rc == 31 and len(res.matches) == 0  # ERROR_INTERNAL_FATAL_ERROR, loudly, no matches


def test_image_layout_matches_header():
    img = open(os.path.join(ROOT, "tests", "golden", "synth_cfg3.img"), "rb").read()
    h = np.frombuffer(img[:48], dtype=np.uint32)
    assert h[0] == 0x31425259 and h[1] == 2
    assert int(np.frombuffer(img[8:16], dtype=np.uint64)[0]) == len(img)
    offs = np.frombuffer(img[48:112], dtype=np.uint64)
    assert all(int(o) % 16 == 0 for o in offs[:6]) and list(offs) == sorted(offs)
