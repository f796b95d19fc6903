"""Seeded differential fuzzing (tests/fuzzgen.py): random text / hex / regex
strings with every modifier, random data with spliced-in instances.
CPU: oracle restatement vs reference.  GPU: drop-in vs reference through the
public API, single block, fast mode, and 1024-byte blocks with overlap."""
import os

import numpy as np
import pytest

import fuzzgen
import yrh

HAVE_REF = os.path.exists(yrh.ref_path())


def _chains(text):
    return "[-]" in text or "-]" in text


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_reference(seed):
    R = yrh.ref()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
    assert n >= 10
    r = R.compile(text)
    data = fuzzgen.gen_data(seed, insts, 1 << 15)
    ref = r.scan(data, cap=1 << 20)
    o = yrh.oracle_scan(r.flatten(), data)
    assert o.error == ref.rc
    if ref.rc == 0:
        got = yrh.matches_from_flat(o.matches, data)
        got["is_private"] = ref.matches["is_private"] if len(got) == len(ref.matches) else 0
        assert got.tobytes() == ref.matches.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_gpu_vs_reference(seed):
    R, B = yrh.ref(), yrh.b200()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 60)
    rr, rb = R.compile(text), B.compile(text)
    data = fuzzgen.gen_data(seed, insts, (1 << 17) + 13 * seed)
    for flags in (0, 1):
        a, b = rr.scan(data, flags=flags, cap=1 << 20), rb.scan(data, flags=flags, cap=1 << 20)
        assert a.rc == b.rc, (a.rc, b.rc)
        if a.rc == 0:
            assert len(a.matches) == len(b.matches)
            assert a.matches.tobytes() == b.matches.tobytes()
            assert a.rule_bits.tobytes() == b.rule_bits.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(100, 108))
def test_gpu_vs_reference_blocks(seed):
    """multi-block scans; chained strings whose parts fall into different blocks included
    (unconfirmed matches are carried across blocks like scanner.c does)"""
    R, B = yrh.ref(), yrh.b200()
    src, insts = fuzzgen.gen_rules(seed, 50)
    good = []
    for s in src:
        try:
            R.compile(s)
            good.append(s)
        except ValueError:
            pass
    text = "".join(good)
    rr, rb = R.compile(text), B.compile(text)
    data = fuzzgen.gen_data(seed, insts, 1 << 14)
    a, b = rr.scan_blocks(data, 1024, 256, cap=1 << 20), rb.scan_blocks(data, 1024, 256, cap=1 << 20)
    assert a.rc == b.rc
    if a.rc == 0:
        assert a.matches.tobytes() == b.matches.tobytes()
        assert a.rule_bits.tobytes() == b.rule_bits.tobytes()
