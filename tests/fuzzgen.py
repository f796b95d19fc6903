"""Seeded random rule + data generator for differential tests (oracle vs
reference on CPU, GPU drop-in vs reference on the B200).  Rules are drawn from a
small grammar over everything the hot path verifies: text strings with every
modifier combination, hex strings with masks / not-bytes / jumps / unbounded
jumps (chains) / alternations, and regular expressions with classes,
quantifiers (greedy and not), groups, alternation, anchors and word boundaries.
Data = random bytes over a small alphabet with instances of the patterns'
literal pieces spliced in, so that the VMs get real work."""
import random

ALPHA = b"abcdxyz019 _.\n\x00\xff"


def _word(rng, lo=2, hi=6):
    return bytes(rng.choice(b"abcdxyz01") for _ in range(rng.randint(lo, hi)))


def gen_text(rng):
    w = _word(rng, 3, 9).decode()
    mods = rng.choice(["", "nocase", "wide", "wide ascii", "nocase wide ascii", "fullword", "fullword nocase",
                       "wide fullword", "xor", "xor(1-40)", "xor wide", "xor wide ascii", "base64", "base64wide", "private"])
    if ("base64" in mods or "xor" in mods) and len(w) < 4:
        w += "abcd"
    return '"%s" %s' % (w, mods), [w.encode()]


def gen_hex(rng):
    toks, pieces, cur = [], [], bytearray()
    n = rng.randint(3, 8)
    for k in range(n):
        r = rng.random()
        last = k == n - 1
        if r < 0.5 or k == 0 or last:
            for _ in range(rng.randint(1, 4)):
                b = rng.choice(ALPHA[:12])
                toks.append("%02X" % b)
                cur.append(b)
        elif r < 0.6:
            toks.append("??")
            cur.append(rng.choice(ALPHA))
        elif r < 0.68:
            b = rng.choice(ALPHA[:12])
            toks.append("%X?" % (b >> 4) if rng.random() < 0.5 else "?%X" % (b & 15))
            cur.append(b)
        elif r < 0.74:
            b = rng.choice(ALPHA[:12])
            toks.append("~%02X" % b)
            cur.append(b ^ 0x55)
        elif r < 0.88:
            a = rng.randint(0, 4)
            bnd = rng.randint(a, a + rng.choice([1, 3, 12]))
            toks.append("[%d-%d]" % (a, bnd))
            cur += bytes(rng.choice(ALPHA) for _ in range(rng.randint(a, bnd)))
        elif r < 0.94:
            toks.append("[-]" if rng.random() < 0.5 else "[%d-]" % rng.randint(1, 3))
            pieces.append(bytes(cur))
            cur = bytearray(bytes(rng.choice(ALPHA) for _ in range(rng.randint(3, 30))))
        else:
            a1, a2 = _word(rng, 1, 3), _word(rng, 1, 3)
            toks.append("( %s | %s )" % (" ".join("%02X" % x for x in a1), " ".join("%02X" % x for x in a2)))
            cur += rng.choice([a1, a2])
    pieces.append(bytes(cur))
    return "{ %s }" % " ".join(toks), [b"".join(pieces)]


def _re_atom(rng, depth=0):
    r = rng.random()
    if r < 0.45:
        w = _word(rng, 1, 4)
        return w.decode(), w
    if r < 0.55:
        return ".", bytes([rng.choice(ALPHA[:12])])
    if r < 0.65:
        cls, inst = rng.choice([("[a-d]", b"b"), ("[^a-c]", b"x"), ("\\d", b"1"), ("\\w", b"z"), ("\\s", b" "), ("\\D", b"y"),
                                ("\\W", b"."), ("\\S", b"a"), ("[0-9a-z]", b"0"), ("[xyz_]", b"_")])
        return cls, inst
    if r < 0.72 and depth < 2:
        a, ia = _re_seq(rng, depth + 1)
        b, ib = _re_seq(rng, depth + 1)
        return "(%s|%s)" % (a, b), rng.choice([ia, ib])
    if r < 0.8 and depth < 2:
        a, ia = _re_seq(rng, depth + 1)
        return "(%s)" % a, ia
    if r < 0.86:
        return rng.choice(["\\b", "\\B"]), b""
    w = _word(rng, 1, 2)
    return w.decode(), w


def _re_seq(rng, depth=0):
    parts, inst = [], b""
    for _ in range(rng.randint(1, 3 if depth else 4)):
        a, ia = _re_atom(rng, depth)
        q = rng.random()
        if a in ("\\b", "\\B"):
            parts.append(a)
            continue
        if q < 0.12:
            a += rng.choice(["*", "*?"]); ia = ia * rng.randint(0, 3)
        elif q < 0.24:
            a += rng.choice(["+", "+?"]); ia = ia * rng.randint(1, 3)
        elif q < 0.32:
            a += rng.choice(["?", "??"]); ia = ia * rng.randint(0, 1)
        elif q < 0.44:
            lo = rng.randint(0, 2); hi = lo + rng.randint(0, 3)
            a += "{%d,%d}%s" % (lo, hi, rng.choice(["", "?"])); ia = ia * rng.randint(lo, hi)
        parts.append(a)
        inst += ia
    return "".join(parts), inst


def gen_regex(rng):
    head = _word(rng, 3, 5)  # a literal head/tail keeps atoms reasonable
    body, inst = _re_seq(rng)
    tail = _word(rng, 2, 4) if rng.random() < 0.7 else b""
    anchor_l = "^" if rng.random() < 0.05 else ""
    anchor_r = "$" if rng.random() < 0.05 else ""
    mods = rng.choice(["", "", "nocase", "wide", "wide ascii", "fullword", "nocase wide"])
    flags = rng.choice(["", "", "s", "i", "is"])
    return "/%s%s%s%s%s/%s %s" % (anchor_l, head.decode(), body, tail.decode(), anchor_r, flags, mods), [head + inst + tail]


def gen_rules(seed, n):
    """-> (source text, list of instance byte strings)"""
    rng = random.Random(seed)
    src, insts = [], []
    k = 0
    while k < n:
        per = rng.randint(1, 6)
        decls = []
        for j in range(per):
            text, ii = rng.choice([gen_text, gen_hex, gen_hex, gen_regex, gen_regex])(rng)
            decls.append("  $s%d = %s" % (j, text))
            insts += ii
        cond = rng.choice(["any of them", "all of them", "#s0 > 1", "$s0 at 0 or any of them", "@s0[1] > 10 or any of them",
                           "!s0[1] > 2 or any of them"])
        src.append("rule r%d {\n strings:\n%s\n condition:\n  %s\n}\n" % (k, "\n".join(decls), cond))
        k += 1
    return src, insts


def gen_data(seed, insts, size):
    rng = random.Random(seed ^ 0xDA7A)
    out = bytearray(rng.choice(ALPHA) for _ in range(size))
    pos = 0
    while pos < size - 64 and insts:
        inst = rng.choice(insts)
        v = rng.random()
        if v < 0.15:
            inst = bytes(c ^ 0x20 if 0x41 <= c <= 0x7a else c for c in inst)     # re-cased
        elif v < 0.3:
            inst = b"".join(bytes((c, 0)) for c in inst)                          # wide
        elif v < 0.36:
            kx = rng.randint(1, 60)
            inst = bytes(c ^ kx for c in inst)                                    # xored
        if pos + len(inst) < size:
            out[pos : pos + len(inst)] = inst
        pos += len(inst) + rng.randint(0, 24)
    return bytes(out)


def compile_filtered(h, seed, n):
    """Compile rule by rule with harness `h`, dropping rules the compiler rejects
    (e.g. a regexp too complex or a hex string the grammar cannot express)."""
    src, insts = gen_rules(seed, n)
    good = []
    for s in src:
        try:
            r = h.compile(s)
            del r
            good.append(s)
        except ValueError:
            pass
    return "".join(good), insts, len(good)
