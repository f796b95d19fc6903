"""Deterministic synthetic rulesets and corpora shaped like BASELINE.json's
configs.  Pure python/numpy; shared by tests/ and bench.py.

Corpus bytes come from a counter-based generator (splitmix64 of the 8-byte word
index) so any byte range of any size can be produced independently on the host
(numpy, here) and on the device (yrb_fill_random in yara_b200/csrc) with
identical contents -- which is what lets byte-range shards and read margins be
generated per rank without moving data.
"""
import random
import string as _string

import numpy as np

WORD = _string.ascii_lowercase
IDENT = _string.ascii_letters + _string.digits + "_"


# ----------------------------------------------------------------- corpus bytes
def rand_bytes(seed: int, start: int, n: int) -> np.ndarray:
    """Bytes [start, start+n) of the infinite stream keyed by `seed`.
    byte k = (splitmix64(seed + (k>>3)) >> (8*(k&7))) & 0xFF."""
    if n <= 0:
        return np.zeros(0, dtype=np.uint8)
    w0 = start >> 3
    w1 = (start + n + 7) >> 3
    with np.errstate(over="ignore"):
        z = np.arange(w0, w1, dtype=np.uint64) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        z = (z + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = z.view(np.uint8)  # little endian: byte j of word = bits 8j..8j+7
    off = start - (w0 << 3)
    return b[off : off + n].copy()


# ------------------------------------------------------------------ string specs
class Spec:
    """One declared string: its YARA source text and a way to make an instance."""

    def __init__(self, text, maker, kind):
        self.text = text      # right-hand side, e.g. '"abc" nocase' or '{ 01 ?? 03 }'
        self.maker = maker    # rng -> bytes
        self.kind = kind

    def instance(self, rng) -> bytes:
        return self.maker(rng)


def _rand_ident(rng, lo=6, hi=23):
    return "".join(rng.choice(IDENT) for _ in range(rng.randint(lo, hi)))


def _rand_word(rng, lo=4, hi=8):
    return "".join(rng.choice(WORD) for _ in range(rng.randint(lo, hi)))


def _wide(b: bytes) -> bytes:
    out = bytearray()
    for c in b:
        out += bytes((c, 0))
    return bytes(out)


def _recase(rng, s: str) -> str:
    return "".join(c.upper() if rng.random() < 0.5 else c.lower() for c in s)


def text_literal(rng) -> Spec:
    s = _rand_ident(rng)
    return Spec('"%s"' % s, lambda r, s=s: s.encode(), "text")


def hex_literal(rng) -> Spec:
    b = bytes(rng.randrange(256) for _ in range(rng.randint(6, 23)))
    return Spec("{ %s }" % " ".join("%02X" % x for x in b), lambda r, b=b: b, "hexlit")


def text_modified(rng) -> Spec:
    s = _rand_ident(rng, 6, 16)
    mod = rng.choice(["nocase", "wide", "wide ascii", "nocase wide", "fullword"])

    def make(r, s=s, mod=mod):
        t = _recase(r, s) if "nocase" in mod else s
        b = t.encode()
        if mod == "wide ascii":
            b = _wide(b) if r.random() < 0.5 else b
        elif "wide" in mod:
            b = _wide(b)
        if mod == "fullword":
            b = b" " + b + b"."
        return b

    return Spec('"%s" %s' % (s, mod), make, "textmod")


def hex_wild(rng, allow_chain=True) -> Spec:
    """Hex string with 1-3 of ??, X?, [a-b] (b<=23), [a-b] with b in 250..1000 (chained)."""
    toks = []  # ('b', v) ('any',) ('hi', v) ('lo', v) ('jump', a, b)

    def lit_run(n):
        for _ in range(n):
            toks.append(("b", rng.randrange(256)))

    lit_run(rng.randint(4, 8))
    kinds = ["any", "nib", "jump"] + (["bigjump"] if allow_chain else [])
    for _ in range(rng.randint(1, 3)):
        k = rng.choice(kinds)
        if k == "any":
            for _ in range(rng.randint(1, 3)):
                toks.append(("any",))
        elif k == "nib":
            toks.append(("hi", rng.randrange(16)) if rng.random() < 0.5 else ("lo", rng.randrange(16)))
        elif k == "jump":
            a = rng.randint(0, 8)
            toks.append(("jump", a, rng.randint(a, 23) if a < 23 else 23))
        else:
            a = rng.randint(0, 300)
            toks.append(("jump", a, rng.randint(max(a, 250), 1000)))
        lit_run(rng.randint(4, 8) if k == "bigjump" else rng.randint(2, 6))

    parts = []
    for t in toks:
        if t[0] == "b":
            parts.append("%02X" % t[1])
        elif t[0] == "any":
            parts.append("??")
        elif t[0] == "hi":
            parts.append("%X?" % t[1])
        elif t[0] == "lo":
            parts.append("?%X" % t[1])
        else:
            parts.append("[%d-%d]" % (t[1], t[2]))

    def make(r, toks=toks):
        out = bytearray()
        for t in toks:
            if t[0] == "b":
                out.append(t[1])
            elif t[0] == "any":
                out.append(r.randrange(256))
            elif t[0] == "hi":
                out.append((t[1] << 4) | r.randrange(16))
            elif t[0] == "lo":
                out.append((r.randrange(16) << 4) | t[1])
            else:
                out += bytes(r.randrange(256) for _ in range(r.randint(t[1], t[2])))
        return bytes(out)

    sp = Spec("{ %s }" % " ".join(parts), make, "hexwild")
    sp.toks = toks
    return sp


def regex(rng) -> Spec:
    k = rng.randrange(5)
    w1, w2, w3 = _rand_word(rng), _rand_word(rng), _rand_word(rng)
    if k == 0:
        text = "/%s[0-9a-f]{4,8}%s/" % (w1, w2)
        make = lambda r: (w1 + "".join(r.choice("0123456789abcdef") for _ in range(r.randint(4, 8))) + w2).encode()
    elif k == 1:
        text = "/%s(%s|%s)\\d+%s/" % (w1, w2, w3, w1[::-1])
        make = lambda r: (w1 + r.choice([w2, w3]) + "".join(r.choice("0123456789") for _ in range(r.randint(1, 6))) + w1[::-1]).encode()
    elif k == 2:
        text = "/%s.{1,16}%s/" % (w1, w2)
        make = lambda r: w1.encode() + bytes(r.choice(b"abcXYZ 0123\t\xff\x80") for _ in range(r.randint(1, 16))) + w2.encode()
    elif k == 3:
        text = "/%s\\s+%s\\w*/ nocase" % (w1, w2)
        make = lambda r: (_recase(r, w1) + " " * r.randint(1, 3) + _recase(r, w2) + "".join(r.choice(IDENT) for _ in range(r.randint(0, 5)))).encode() + b"."
    else:
        text = "/\\b%s[a-z]*%s\\b/" % (w1, w2)
        make = lambda r: b" " + (w1 + "".join(r.choice(WORD) for _ in range(r.randint(0, 6))) + w2).encode() + b" "
    return Spec(text, make, "regex")


_B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"


def _b64_variant(s: bytes, phase: int) -> bytes:
    """What `"s" base64` can look like inside a base64 stream when s starts at
    byte `phase` of a 3-byte group (base64.c:383-482 generates the 3 variants)."""
    import base64 as _b

    enc = _b.b64encode(b"\x00" * phase + s + b"\x00\x00").decode()
    # drop the leading chars polluted by the pad prefix and the trailing ones polluted by the suffix
    lead = {0: 0, 1: 2, 2: 3}[phase]
    total_bits = (phase + len(s)) * 8
    full = total_bits // 6
    return enc[lead:full].encode()


def text_special(rng) -> Spec:
    s = _rand_ident(rng, 8, 20)
    mod = rng.choice(["xor", "xor(1-32)", "xor wide", "base64", "base64wide"])

    def make(r, s=s, mod=mod):
        b = s.encode()
        if mod.startswith("xor"):
            k = r.randint(1, 32) if "1-32" in mod else r.randrange(256)
            if "wide" in mod:
                b = _wide(b)
            return bytes(x ^ k for x in b)
        v = _b64_variant(b, r.randrange(3))
        return _wide(v) if mod == "base64wide" else v

    return Spec('"%s" %s' % (s, mod), make, "special")


# ---------------------------------------------------------------------- rulesets
class Ruleset:
    def __init__(self, name, specs, per_rule=50):
        self.name = name
        self.specs = specs
        self.per_rule = per_rule
        self.extra = ""   # rule text appended verbatim (bench.py --short-atoms)

    @property
    def source(self) -> str:
        out = []
        for r0 in range(0, len(self.specs), self.per_rule):
            chunk = self.specs[r0 : r0 + self.per_rule]
            out.append("rule r%d {\n strings:\n" % (r0 // self.per_rule))
            for j, sp in enumerate(chunk):
                out.append("  $s%d = %s\n" % (j, sp.text))
            out.append(" condition:\n  any of them\n}\n")
        return "".join(out) + self.extra


def ruleset(config: int, n: int = None, seed: int = None) -> Ruleset:
    """config 2: 1k literals | 3: 20k hex-wild/jump + text modifiers |
    4: 100k mixed | 5: 5k regex + xor/base64.  `n` overrides the declared count."""
    defaults = {2: (1000, 1), 3: (20000, 2), 4: (100000, 3), 5: (5000, 4)}
    dn, dseed = defaults[config]
    n = dn if n is None else n
    rng = random.Random(dseed if seed is None else seed)
    specs = []
    for i in range(n):
        if config == 2:
            specs.append(text_literal(rng) if i % 2 == 0 else hex_literal(rng))
        elif config == 3:
            specs.append(hex_wild(rng) if i % 3 != 2 else text_modified(rng))
        elif config == 4:
            specs.append([text_literal, hex_literal, lambda r: hex_wild(r), text_modified, regex][i % 5](rng))
        else:
            specs.append(regex(rng) if i % 6 != 5 else text_special(rng))
    return Ruleset("config%d_n%d" % (config, n), specs)


# ----------------------------------------------------------------------- corpora
def plants(rs: Ruleset, size: int, every: int, seed: int = 11, start: int = 0, edge=True):
    """Deterministic plant list [(offset, bytes)] for the byte range
    [start, start+size) of a corpus: about one instance per `every` bytes, never
    overlapping, each fully inside its own `every`-sized cell -- so a shard can
    compute the plants of its own range (and margins) without seeing the rest."""
    out = []
    c0 = start // every
    c1 = (start + size + every - 1) // every
    for cell in range(c0, c1):
        rng = random.Random((seed << 40) ^ cell)
        sp = rs.specs[rng.randrange(len(rs.specs))]
        inst = sp.instance(rng)
        if len(inst) + 2 > every:
            continue
        off = cell * every + rng.randrange(every - len(inst))
        out.append((off, inst))
    if edge and start == 0 and rs.specs:
        rng = random.Random(seed ^ 0x5EED)
        out.append((0, rs.specs[0].instance(rng)))          # starts at offset 0
    return out


def _chained_instance(sp: Spec, rng, min_gap=0):
    """An instance of a hex string with a big jump ([a-b], b >= 250: the compiler splits it into a chain,
    re.c:452-512) whose first big jump is at least min_gap bytes long; returns the bytes and the offset of
    the middle of that jump."""
    out = bytearray()
    cut = None
    for t in sp.toks:
        if t[0] == "b":
            out.append(t[1])
        elif t[0] == "any":
            out.append(rng.randrange(256))
        elif t[0] == "hi":
            out.append((t[1] << 4) | rng.randrange(16))
        elif t[0] == "lo":
            out.append((rng.randrange(16) << 4) | t[1])
        else:
            n = rng.randint(t[1], t[2])
            if cut is None and t[2] >= 250:
                n = min(max(n, min_gap), t[2])
                cut = len(out) + n // 2
            out += bytes(rng.randrange(256) for _ in range(n))
    return bytes(out), cut


def boundary_plants(rs: Ruleset, block_size: int, parts: int = 8, seed: int = 23):
    """an instance straddling EVERY shard boundary, and chained strings whose parts
    lie in different shards.  Boundaries are those of yrb_multi_plan / shard.plan for `parts` shards (which
    include those for every divisor of `parts`): multiples of ceil(block / parts) rounded up to 4096.  At
    each boundary b: a chained instance placed so that b falls into the middle of its big jump (head part
    in one shard, tail part in the next), and -- written after it, into the jump's filler bytes -- an
    instance of an ordinary string cut by b at a random byte.  Deterministic, and independent of how many
    shards actually scan the block.  Apply AFTER plants(): these overwrite whatever the cells put there."""
    out = []
    per = ((block_size + parts - 1) // parts + 4095) & ~4095
    chained = [sp for sp in rs.specs if getattr(sp, "toks", None) and any(t[0] == "jump" and t[2] >= 250 for t in sp.toks)]
    plain = [sp for sp in rs.specs if sp.kind != "regex" and sp not in chained] or rs.specs
    for k in range(1, parts):
        b = k * per
        if b + 8192 >= block_size:
            continue
        rng = random.Random((seed << 32) ^ k)
        inst = plain[rng.randrange(len(plain))].instance(rng)
        if chained:
            cinst, cut = _chained_instance(chained[rng.randrange(len(chained))], rng, min_gap=len(inst) + 16)
            if cut is not None and b - cut >= 0:
                out.append((b - cut, cinst))
        out.append((b - rng.randint(1, len(inst) - 1), inst))
    return out


def corpus(rs: Ruleset, size: int, every: int = 65536, seed: int = 7, plant_seed: int = 11, tail_plant=True):
    """Host corpus of `size` bytes: counter-based random bytes with planted instances."""
    buf = rand_bytes(seed, 0, size)
    pl = plants(rs, size, every, plant_seed) if every else []
    for off, inst in pl:
        if off + len(inst) <= size:
            buf[off : off + len(inst)] = np.frombuffer(inst, dtype=np.uint8)
    if tail_plant and rs.specs and every:
        rng = random.Random(plant_seed ^ 0x7A11)
        sp = rs.specs[-1 if rs.specs[-1].kind != "regex" else 0]
        inst = sp.instance(rng)
        if len(inst) < size:
            buf[size - len(inst) :] = np.frombuffer(inst, dtype=np.uint8)  # ends at the last byte
        # overlapping self-repeat and two instances one byte apart
        sp0 = rs.specs[0]
        i0 = sp0.instance(rng)
        mid = size // 2
        if mid + 2 * len(i0) + 2 < size and mid > 0:
            buf[mid : mid + len(i0)] = np.frombuffer(i0, dtype=np.uint8)
            buf[mid + len(i0) + 1 : mid + 2 * len(i0) + 1] = np.frombuffer(i0, dtype=np.uint8)
    return buf
