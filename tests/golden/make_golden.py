#!/usr/bin/env python3
"""Generates the golden fixtures of tests/golden/ from the REFERENCE libyara
(oracle/_ref/libyara_ref.so, built from /root/reference by oracle/Makefile).
Run here (the reference does not exist on the GPU box); commit the outputs.

Each fixture = flat rules image (.img) + (string_idx, offset, match_length,
xor_key) lists the reference produced for a corpus that is either stored (.bin)
or regenerated from yara_b200.synth seeds."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import yrh  # noqa: E402
from cases import CASES  # noqa: E402
from yara_b200 import synth  # noqa: E402


def dump(name, rules, data, meta, flags=0):
    ref = rules.scan(data, flags=flags, cap=1 << 18)
    open(os.path.join(HERE, name + ".img"), "wb").write(rules.flatten())
    meta = dict(meta)
    meta.update(image=name + ".img", rc=ref.rc, flags=flags,
                matches=[[int(m["string_idx"]), int(m["offset"]), int(m["match_length"]), int(m["xor_key"])]
                         for m in ref.matches])
    json.dump(meta, open(os.path.join(HERE, name + ".json"), "w"))
    print(name, len(ref.matches), "matches")


def main():
    R = yrh.ref()
    # the quirk-bearing known-answer cases
    only = [int(x) for x in sys.argv[1:] if x.isdigit()]   # `make_golden.py 59`: (re)generate just these cases
    for k in (only or (8, 19, 22, 27, 28, 32, 33, 42, len(CASES) - 1)):
        rule, data = CASES[k]
        open(os.path.join(HERE, "case%02d.bin" % k), "wb").write(data)
        dump("case%02d" % k, R.compile(rule), data, {"data": "case%02d.bin" % k, "rule": rule})
    if only:
        return
    # match chains of 300 entries (tests/long_chains.py): more than an 8-bit chain position holds
    import long_chains
    data = long_chains.corpus(4096, seed=3)
    open(os.path.join(HERE, "long_chains.bin"), "wb").write(data.tobytes())
    dump("long_chains", R.compile(long_chains.ruleset(600)), data, {"data": "long_chains.bin", "rules": "long_chains.ruleset(600)"})
    if "long_chains" in sys.argv:
        return
    for cfg, n, size, every in ((2, 200, 1 << 16, 1024), (3, 300, 1 << 16, 512), (5, 120, 1 << 16, 512)):
        rs = synth.ruleset(cfg, n)
        dump("synth_cfg%d" % cfg, R.compile(rs.source), synth.corpus(rs, size, every),
             {"config": cfg, "n": n, "size": size, "every": every})


if __name__ == "__main__":
    main()
