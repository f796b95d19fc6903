"""CPU tests of the condition-VM pre-filter (yara_b200/host/yr_b200_prefilter.c).

oracle/_ref/libyara_refpf.so is the reference library with ONLY the pre-filter hooks applied
(exec.c OP_INIT_RULE / OP_MATCH_RULE, begin/end around yr_execute_code, matches noted in
_yr_scan_match_callback) -- the same yr_b200_prefilter.c and the same exec.c edits the GPU drop-in is
built with.  With the filter ON it must be indistinguishable from the untouched reference: same rule
verdicts and matches on every scan, including the scans AFTER the filter has learnt which rules it may
skip."""
import os
import random
import subprocess

import numpy as np
import pytest

os.environ["YR_B200_RULE_PREFILTER"] = "1"   # (the default) read once, at the first scan of the refpf library

import fuzzgen
import yrh
from cases import CASES
from yara_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
HAVE = os.path.exists(yrh.refpf_path()) and os.path.exists(yrh.ref_path())
pytestmark = pytest.mark.skipif(not HAVE, reason="oracle/_ref not built (needs /root/reference at build time)")


def same(a, b):
    assert a.rc == b.rc
    assert a.rule_bits.tobytes() == b.rule_bits.tobytes()
    assert a.matches.tobytes() == b.matches.tobytes()


CONDITIONS = [
    "$a", "not $a", "$a and $b", "$a or $b", "any of them", "all of them", "none of them", "1 of them", "2 of them",
    "0 of them", "#a == 0", "#a > 1", "#a + #b >= 2", "@a[1] < 10", "!a[1] == 6", "$a at 0", "$a in (0..5)",
    "for any of them : ( $ )", "for all of them : ( $ )", "for all of them : ( # > 1 )", "for any i in (1..#a) : ( @a[i] > 3 )",
    "for all i in (1, 2, 3) : ( i < #b + 4 )", "50% of them", "any of ($a*) in (0..100)", "#a in (0..20) == 1",
    "defined @a[3]", "not defined @a[1]", "true", "false", "$a and true", "not $b or #a > 5",
    # impure ones: data, file size, other rules
    "filesize > 3 and $a", "filesize < 3", "uint8(0) == 0x6e or $a", "$a and uint16(1) != 7", "entrypoint == 0 or $b",
]
BUFFERS = [b"nothing here", b"needle", b"nothing here", b"haystack needle", b"xx needle needle", b"", b"haystack", b"nothing here",
           b"needle haystack needle", b"n"]


def test_every_condition_shape_stays_identical_over_repeated_scans():
    import re

    def decl(c):
        them = "them" in c
        out = ""
        if them or re.search(r"[$#@!]a", c):
            out += ' $a = "needle"'
        if them or re.search(r"[$#@!]b", c):
            out += ' $b = "haystack"'
        return "strings:" + out if out else ""

    rules = "".join("rule r%d { %s condition: %s }\n" % (k, decl(c), c) for k, c in enumerate(CONDITIONS))
    rules += 'global rule g { strings: $a = "hay" condition: $a or filesize < 100 }\n'
    rules += 'private rule p { strings: $a = "needle" condition: $a }\n'
    rules += 'rule uses_p { strings: $b = "stack" condition: p and $b }\n'
    rules += 'rule of_rules { condition: any of (r0, r3) }\n'
    R, P = yrh.ref(), yrh.refpf()
    rr, rp = R.compile(rules), P.compile(rules)
    e0, s0 = yrh.prefilter_stats(P)
    for rounds in range(3):
        for buf in BUFFERS:
            for flags in (0, 1):
                same(rr.scan(buf, flags=flags), rp.scan(buf, flags=flags))
    e1, s1 = yrh.prefilter_stats(P)
    cls = yrh.prefilter_classes(P, rp)
    assert s1 > s0, "nothing was ever skipped"
    assert cls["walked"] == rp.num_rules          # the classifier decoded the whole condition code
    assert cls["skippable"] >= 15 and cls["true_without_matches"] >= 5
    # data / filesize / rule references / the global rule must never be classified pure
    assert cls["impure"] >= 5 + 1 + 2


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_known_answer_cases(idx):
    rule, data = CASES[idx]
    R, P = yrh.ref(), yrh.refpf()
    rr, rp = R.compile(rule), P.compile(rule)
    noise = bytes(len(data) if len(data) < 64 else 64)
    for buf in (noise, data, noise, data, b"", data):
        same(rr.scan(buf), rp.scan(buf))
    same(rr.scan_blocks(data, 1024, 256), rp.scan_blocks(data, 1024, 256))
    assert yrh.prefilter_classes(P, rp)["walked"] == rp.num_rules


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_rulesets(seed):
    R, P = yrh.ref(), yrh.refpf()
    text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
    rr, rp = R.compile(text), P.compile(text)
    data = fuzzgen.gen_data(seed, insts, 1 << 14)
    noise = synth.rand_bytes(seed, 0, 4096).copy()
    half = np.ascontiguousarray(data[: len(data) // 2])
    for buf in (noise, data, noise, half, data):
        for flags in (0, 1):
            same(rr.scan(buf, flags=flags, cap=1 << 20), rp.scan(buf, flags=flags, cap=1 << 20))


def test_file_batch_shape_skips_almost_every_rule():
    """BASELINE configs[3] shape: after the first file, a rule is evaluated only when one of its strings matched"""
    rs = synth.ruleset(4, 3000)
    R, P = yrh.ref(), yrh.refpf()
    rr, rp = R.compile(rs.source), P.compile(rs.source)
    files = [synth.corpus(rs, 20000 + 1000 * k, 4096, seed=k, plant_seed=70 + k) for k in range(12)]
    same(rr.scan(files[0]), rp.scan(files[0]))
    e0, s0 = yrh.prefilter_stats(P)
    for f in files[1:]:
        same(rr.scan(f), rp.scan(f))
    e1, s1 = yrh.prefilter_stats(P)
    nrules, nfiles = rp.num_rules, len(files) - 1
    assert (e1 - e0) + (s1 - s0) == nrules * nfiles
    assert (s1 - s0) > 0.5 * nrules * nfiles
    cls = yrh.prefilter_classes(P, rp)
    assert cls["impure"] == 0 and cls["walked"] == nrules


@pytest.mark.parametrize("name", ["test-rules_refpf", "test-api_refpf", "test-async_refpf"])
def test_reference_programs_with_the_filter_on(name):
    """the reference's own test programs, compiled unmodified, against the hooked library"""
    exe = os.path.join(REFDIR, name)
    if not os.path.exists(exe):
        pytest.skip("%s not built" % name)
    env = dict(os.environ, TOP_SRCDIR=os.path.join(REFDIR, "top"), YR_B200_RULE_PREFILTER="1")
    p = subprocess.run([exe], cwd=REFDIR, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]


def test_threads_sharing_one_rules_object():
    """yr_rules_scan_mem from several threads over one YR_RULES (limits.h YR_MAX_THREADS): the learnt
    classification is shared, the per-scan match marks are not"""
    from concurrent.futures import ThreadPoolExecutor
    rs = synth.ruleset(4, 2000)
    R, P = yrh.ref(), yrh.refpf()
    rr, rp = R.compile(rs.source), P.compile(rs.source)
    files = [synth.corpus(rs, 30000 + 777 * k, 2048, seed=k, plant_seed=90 + k) for k in range(48)]
    want = [rr.scan(f).key() for f in files]
    for _ in range(2):
        with ThreadPoolExecutor(8) as ex:
            got = list(ex.map(lambda f: rp.scan(f).key(), files))
        assert got == want


def test_many_threads_learning_at_once():
    """OP_INIT_RULE asks the pre-filter twice per rule (the jump, and the operand stepped over when the jump is not
    taken); the learnt table is shared by every thread scanning with the same rules, so another thread can turn a rule
    "skippable" between the two questions.  The answers must agree (found with `yara -p 16` over the 100,000-string
    set through the drop-in: the first scans of some threads ended in ERROR_COULD_NOT_MAP_FILE, the exception
    handler's code).  3,000 pure rules, 16 threads that all start learning at the same time, fresh rules each round."""
    from yara_b200 import synth
    P = yrh.refpf()
    for k in range(4):
        rs = synth.ruleset(2, 6000, seed=100 + k)
        rs.per_rule = 2
        r = P.compile(rs.source)
        data = synth.corpus(rs, 1 << 20, 4096, seed=k)
        secs, rc = r.scan_mt(data, 16)
        assert rc == 0, (k, rc)


def test_one_long_lived_scanner_and_odd_block_lists():
    """the usage patterns of tests/test_glue_cpu.py -- ONE YR_SCANNER over many scans of different kinds, block lists
    of any shape, a scanner reused after a scan that timed out -- with the filter on: what it has learnt in one
    scan is used in the next ones of the same scanner, and nothing observable changes"""
    import test_glue_cpu as t
    P = yrh.refpf()
    t.run_scanner_sequence(P, range(5))
    t.run_regions(P, range(2))
    t.run_scanner_reuse_after_timed_out_scan(P)
