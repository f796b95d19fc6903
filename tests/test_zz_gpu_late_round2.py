"""GPU parity (-m gpu): the tests added at the end of round 2, after the GPU budget of the round was spent on the
others (the file sorts last on purpose; what ran on a B200 and what only on the CPU stand-in is in profiles/README.md).

* Rulesets with LONG YR_AC_MATCH chains (tests/long_chains.py: hundreds of rules that share one string hang hundreds of
  entries on one Aho-Corasick state).  The discovery key gives the chain position as many bits as the ruleset's
  longest chain needs (engine.cuh: DevRules::key_seq_bits); chains of more than 255 entries used to be refused at
  upload.  Through the drop-in against the reference library (whole block in both modes, 1024-byte blocks, a parked
  file batch) and through the engine with both AC kernels against the oracle.
* A reused YR_SCANNER after a multi-block scan that ended early (yr_b200_blocks_discard).
* Fast mode over several blocks."""
import numpy as np
import pytest

import long_chains
import yrh
from test_glue_cpu import run_long_chains

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [600, 1200])
def test_long_match_chains_through_the_dropin(n):
    assert run_long_chains(yrh.b200(), n) > 10 * n


def test_long_match_chains_both_ac_kernels():
    from yara_b200 import api
    from yara_b200.api import Engine
    from test_vm_host_cpu import canon
    fields = ("string_idx", "offset", "match_length", "xor_key")
    image = yrh.ref().compile(long_chains.ruleset(1200)).flatten()
    data = long_chains.corpus(1 << 18, seed=9)
    o = yrh.oracle_scan(image, data)
    assert o.error == 0 and len(o.matches) > 100000
    want = canon(o.matches, fields)
    eng = Engine(image, 0)
    for kernel in (api.KERNEL_WALK, api.KERNEL_FILTER):       # FILTER: the filter kernel if the ruleset admits it
        eng.set_option(api.OPT_SCAN_KERNEL, kernel)
        r = eng.scan_host(data)
        assert r.rc == 0 and r.error == 0
        assert canon(r.matches, fields) == want
    eng.close()


def test_scanner_reuse_after_a_scan_that_ended_early():
    """tests/test_glue_cpu.py: blocks deferred by a multi-block scan that ran out of time are dropped with the scan
    (yr_b200_blocks_discard at scanner.c:545), not scanned as part of the scanner's next scan."""
    from test_glue_cpu import run_scanner_reuse_after_timed_out_scan
    run_scanner_reuse_after_timed_out_scan(yrh.b200())


def test_fast_mode_over_several_blocks():
    """SCAN_FLAGS_FAST_MODE with 1024-byte blocks (the one combination of the reference's scan modes the GPU suite did
    not have): known-answer cases and fuzz rulesets against the reference library.  A chained string's head that
    already has a match is not verified again in later blocks, but chains a later block's tail completes -- also
    with head matches an earlier block left unconfirmed -- still count (yr_b200_glue.c scan_block_now)."""
    import fuzzgen
    from cases import CASES
    R, B = yrh.ref(), yrh.b200()
    for idx, (rule, data) in enumerate(CASES):
        a = R.compile(rule).scan_blocks(data, 1024, 256, flags=1, cap=1 << 20)
        b = B.compile(rule).scan_blocks(data, 1024, 256, flags=1, cap=1 << 20)
        assert a.rc == b.rc, (idx, a.rc, b.rc)
        if a.rc == 0:
            assert a.matches.tobytes() == b.matches.tobytes(), idx
            assert a.rule_bits.tobytes() == b.rule_bits.tobytes(), idx
    for seed in (0, 1, 2, 3, 100, 101, 102, 103):
        text, insts, n = fuzzgen.compile_filtered(R, seed, 40)
        rr, rb = R.compile(text), B.compile(text)
        data = np.frombuffer(bytes(fuzzgen.gen_data(seed + 300, insts, (1 << 13) + 291)), dtype=np.uint8)
        a, b = rr.scan_blocks(data, 1024, 256, flags=1, cap=1 << 20), rb.scan_blocks(data, 1024, 256, flags=1, cap=1 << 20)
        assert a.rc == b.rc, (seed, a.rc, b.rc)
        if a.rc == 0:
            assert a.matches.tobytes() == b.matches.tobytes(), seed
            assert a.rule_bits.tobytes() == b.rule_bits.tobytes(), seed
