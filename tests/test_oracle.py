"""CPU tests (no GPU): the oracle restatement (oracle/yr_oracle.c over the flat
rules image) against the reference libyara built from /root/reference
(oracle/_ref) -- known-answer cases and seeded synthetic rulesets -- and against
the committed golden fixtures (tests/golden/, generated from the reference by
tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

import yrh
from cases import CASES
from yara_b200 import synth

HAVE_REF = os.path.exists(yrh.ref_path())
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def compare(ref, o, data, rules):
    got = yrh.matches_from_flat(o.matches, data)
    # is_private is a property of the string, not of the scan
    priv = {}
    for m in ref.matches:
        priv[int(m["string_idx"])] = int(m["is_private"])
    for i in range(len(got)):
        got["is_private"][i] = priv.get(int(got["string_idx"][i]), 0)
    assert o.error == ref.rc
    assert len(got) == len(ref.matches)
    assert got.tobytes() == ref.matches.tobytes()


@needs_ref
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases(idx):
    rule, data = CASES[idx]
    r = yrh.ref().compile(rule)
    ref = r.scan(data)
    o = yrh.oracle_scan(r.flatten(), data)
    if ref.rc != 0:     # cases that pin an error code (a failed scan reports no matches)
        assert o.error == ref.rc
        return
    compare(ref, o, data, r)


@needs_ref
@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases_fast_mode(idx):
    rule, data = CASES[idx]
    r = yrh.ref().compile(rule)
    ref = r.scan(data, flags=1)
    o = yrh.oracle_scan(r.flatten(), data, flags=1)
    compare(ref, o, data, r)


@needs_ref
@pytest.mark.parametrize("cfg,n,size,every", [(2, 1000, 1 << 21, 4096), (3, 2000, 1 << 21, 2048),
                                              (4, 3000, 1 << 21, 2048), (5, 500, 1 << 20, 1024)])
def test_synthetic_configs(cfg, n, size, every):
    rs = synth.ruleset(cfg, n)
    r = yrh.ref().compile(rs.source)
    data = synth.corpus(rs, size, every)
    ref = r.scan(data, cap=1 << 18)
    assert ref.rc == 0 and len(ref.matches) > size // every // 2
    o = yrh.oracle_scan(r.flatten(), data)
    compare(ref, o, data, r)


@needs_ref
def test_too_many_matches_cap():
    """test-api.c:320-395: 2M x 'a' vs "aa": 1,000,000 matches kept, one TOO_MANY callback."""
    r = yrh.ref().compile('rule t { strings: $a = "aa" condition: $a }')
    data = np.full(1_200_000, ord("a"), dtype=np.uint8)
    ref = r.scan(data, cap=1_100_000)
    assert ref.rc == 0 and ref.too_many_calls == 1 and len(ref.matches) == 1_000_000
    o = yrh.oracle_scan(r.flatten(), data)
    assert o.error == 0 and o.too_many == [0] and len(o.matches) == 1_000_000
    ref2 = r.scan(data, too_many_action=1, cap=16)
    o2 = yrh.oracle_scan(r.flatten(), data, too_many_action=1)
    assert ref2.rc == 30 and o2.error == 30


def test_golden_fixtures():
    """Golden vectors generated from the reference build (make_golden.py) -- usable
    without oracle/_ref."""
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.json")))
    assert files, "no golden fixtures"
    for f in files:
        meta = json.load(open(f))
        img = open(os.path.join(GOLDEN, meta["image"]), "rb").read()
        data = np.fromfile(os.path.join(GOLDEN, meta["data"]), dtype=np.uint8) if meta.get("data") else \
            synth.corpus(synth.ruleset(meta["config"], meta["n"]), meta["size"], meta["every"])
        o = yrh.oracle_scan(img, data, flags=meta.get("flags", 0))
        got = [(int(m["string_idx"]), int(m["offset"]), int(m["match_length"]), int(m["xor_key"])) for m in o.matches]
        want = [tuple(x) for x in meta["matches"]]
        assert o.error == meta["rc"]
        assert got == want, f
