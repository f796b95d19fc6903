"""Rulesets whose Aho-Corasick states carry LONG YR_AC_MATCH chains: many rules that share a string share its
atom, and every one of them hangs an entry on the same state's chain (ahocorasick.c: _yr_ac_add_string /
_yr_ac_create_failure_links).  Real rule collections do this all the time (the same API name or command line in
hundreds of rules); the engine's discovery key has to order all the entries of such a chain (engine.cuh:
DevRules::key_seq_bits), not only the first 255.  Used by the CPU replay, the glue and the GPU parity tests."""
import numpy as np

KINDS = [
    '$a = "kernel32.dll"',
    '$a = "kernel32.dll" fullword',
    '$a = /kernel32\\.dl+/',
    '$a = { 6B 65 72 6E 65 6C 33 32 [1-300] 6B 65 72 6E 65 6C }',      # a chain: head and tail share the atom
    '$a = { 6B 65 72 6E 65 6C ?? 32 2E 64 6C 6C }',
    '$a = "kernel32.dll" private',
    '$a = /kernel32(\\.dll|\\.sys)?/',
    '$a = "kernel32"',
]


def ruleset(n):
    return "".join("rule r%d { strings: %s condition: $a }\n" % (i, KINDS[i % len(KINDS)]) for i in range(n))


def corpus(size=1 << 15, seed=5):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size, dtype=np.uint8)
    words = [b"kernel32.dll", b" kernel32.dll ", b"kernel32.dllll", b"kernel32.sys", b"xkernel32.dllx", b"kernel32",
             b"kernel32.dll" + b"A" * 40 + b"kernel", b"kernelX2.dll", b"kernel32.dl"]
    pos = 64
    k = 0
    while pos + 400 < size:
        w = words[k % len(words)]
        d[pos:pos + len(w)] = np.frombuffer(w, dtype=np.uint8)
        pos += len(w) + int(rng.integers(1, 350))
        k += 1
    d[:12] = np.frombuffer(b"kernel32.dll", dtype=np.uint8)          # at offset 0 and flush with the end
    d[size - 12:] = np.frombuffer(b"kernel32.dll", dtype=np.uint8)
    return d
