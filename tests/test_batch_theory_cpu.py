"""CPU check of the argument yrb_scan_batch rests on (yara_b200/csrc/emit.cuh):

    walking the CONCATENATION of independent items and applying the reference's guard
    `backtrack <= i` (scanner.c:104) with i counted from the start of the item that holds the
    last consumed byte yields, item by item and in the same order, exactly the
    yr_scan_verify_match calls of walking every item on its own from the root state.

Just behind an item boundary the automaton may be in a deeper state than a fresh walk would be;
the entries that state adds in front of the shared tail of its match list all have
backtrack >= depth > bytes consumed in the item, so the guard drops them.  Checked here with the
oracle's Aho-Corasick walk (oracle/yr_oracle.c ac_walk) on fuzz and synthetic rulesets, with
adversarial gap bytes between the items."""
import os
import random

import numpy as np
import pytest

import fuzzgen
import yrh
from yara_b200 import synth

HAVE_REF = os.path.exists(yrh.ref_path())
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")


def acm_backtracks(image: bytes):
    num_acm = int(np.frombuffer(image[20:24], dtype=np.uint32)[0])
    off_acm = int(np.frombuffer(image[64:72], dtype=np.uint64)[0])
    a = np.frombuffer(image, dtype=np.uint32, count=num_acm * 5, offset=off_acm).reshape(num_acm, 5)
    return a[:, 4].astype(np.int64)


def root_has_matches(image: bytes):
    off_match = int(np.frombuffer(image[56:64], dtype=np.uint64)[0])
    return int(np.frombuffer(image, dtype=np.uint32, count=1, offset=off_match)[0]) != 0


def check(image, items, rng):
    if root_has_matches(image):
        pytest.skip("ruleset with root-state matches: batches are refused for those")
    bt = acm_backtracks(image)
    starts, parts, cur = [], [], 0
    for it in items:
        starts.append(cur)
        parts.append(it)
        cur += len(it)
        gap = rng.choice([0, 0, 1, 3, 15])
        if gap:
            # gap bytes copied from the data itself: as atom-like as it gets
            src = items[rng.randrange(len(items))]
            g = np.resize(src, gap) if len(src) else np.zeros(gap, np.uint8)
            parts.append(g.astype(np.uint8))
            cur += gap
    concat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    pos, acm = yrh.oracle_ac_events(image, concat)
    starts_a = np.array(starts, dtype=np.int64)
    lens_a = np.array([len(i) for i in items], dtype=np.int64)

    got = [[] for _ in items]
    for i, a in zip(pos.astype(np.int64), acm):
        if i == 0:
            continue
        last = i - 1
        k = int(np.searchsorted(starts_a, last, side="right")) - 1
        # zero-length items share their start with the next one: the owner is the last item starting there
        if k < 0 or last >= starts_a[k] + lens_a[k]:
            continue                      # a gap byte
        rel = i - starts_a[k]
        if bt[a] <= rel:
            got[k].append((int(rel), int(a)))

    total = 0
    for k, it in enumerate(items):
        p, a = yrh.oracle_ac_events(image, it)
        want = [(int(x), int(y)) for x, y in zip(p, a)]
        assert got[k] == want, "item %d of %d" % (k, len(items))
        total += len(want)
    return total


def ragged_items(data, rng, n):
    cuts = sorted(set([0, len(data)] + [rng.randrange(len(data) + 1) for _ in range(n)]))
    items = [np.ascontiguousarray(data[a:b]) for a, b in zip(cuts, cuts[1:])]
    items.insert(rng.randrange(len(items) + 1), np.zeros(0, np.uint8))
    items.append(np.ascontiguousarray(data[:3]))
    return items


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_rulesets(seed):
    R = yrh.ref()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
    image = R.compile(text).flatten()
    data = np.frombuffer(bytes(fuzzgen.gen_data(seed, insts, 1 << 14)), dtype=np.uint8)
    rng = random.Random(seed)
    assert check(image, ragged_items(data, rng, 40), rng) > 0


@pytest.mark.parametrize("cfg,n", [(2, 300), (3, 1000), (4, 1500), (5, 400)])
def test_synthetic_rulesets(cfg, n):
    rs = synth.ruleset(cfg, n)
    image = yrh.ref().compile(rs.source).flatten()
    data = synth.corpus(rs, 1 << 17, 512)
    rng = random.Random(cfg)
    assert check(image, ragged_items(data, rng, 200), rng) > 100
