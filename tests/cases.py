"""Known-answer style cases shared by the CPU (oracle) and GPU (engine) parity
tests.  Rules/data mirror the shapes tests/test-rules.c of the reference pins
(test_strings :613-1279, test_hex_strings :1300-1624, test_re :2252-2734) plus
the quirks listed in SURVEY.md 8(a)."""

T1024 = (b"abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ\n" * 20)[:1024]
BLOB = T1024 + b"---- abc ---- xyz" + b"\x00" * 7 + b"---- a\x00b\x00c\x00 ---- x\x00y\x00z\x00"


def R(strings, cond="any of them"):
    return "rule t { strings: %s condition: %s }" % (strings, cond)


CASES = [
    # literals, modifiers (test_strings)
    (R('$a = "abc"'), b"---- abc ---- xyz abcabc"),
    (R('$a = "abc" nocase wide ascii'), b"---- ABC ---- a\0b\0c\0 A\0b\0C\0abcabc"),
    (R('$a = "abc" fullword'), b"abc xabc abc. abcabc abc"),
    (R('$a = "abc" wide fullword'), b"a\0b\0c\0 x\0a\0b\0c\0 .\0a\0b\0c\0.\0 a\0b\0c\0"),
    (R('$a = "abc" $b = "xyz" $c = "missing"', "$a and $b and not $c"), BLOB),
    (R('$a = "abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"'), T1024),
    (R('$a = "ab"'), b"xxabab ab"),
    (R('$a = "a"'), b"banana"),
    (R('$a = "abcd" xor'), bytes(b ^ 0x55 for b in b"abcd") + b"abcd"),              # quirk 2: FITS_IN_ATOM xor key stays 0
    (R('$a = "hello world" xor'), bytes(b ^ 0x55 for b in b"hello world") + b"hello world"),
    (R('$a = "hello world" xor(1-32)'), bytes(b ^ 0x7 for b in b"hello world") + bytes(b ^ 0x55 for b in b"hello world")),
    (R('$a = "hello" xor wide'), bytes(b ^ 0x21 for b in b"h\0e\0l\0l\0o\0") + b"h\0e\0l\0l\0o\0"),
    (R('$a = "hello" xor wide ascii'), bytes(b ^ 0x21 for b in b"h\0e\0l\0l\0o\0") + bytes(b ^ 3 for b in b"hello")),
    (R('$a = "This program cannot" base64'), b"VGhpcyBwcm9ncmFtIGNhbm5vdA== RoaXMgcHJvZ3JhbSBjYW5ub3Q UaGlzIHByb2dyYW0gY2Fubm90"),
    (R('$a = "This program cannot" base64wide'), b"V\0G\0h\0p\0c\0y\0B\0w\0c\0m\09\0n\0c\0m\0F\0t\0I\0G\0N\0h\0b\0m\05\0v\0d\0A\0"),
    (R('$a = "abc" private $b = "xyz"'), BLOB),
    (R('$a = "abc"', "$a at 1029"), BLOB),
    (R('$a = "abc"', "$a at 0"), b"abc abc"),
    (R('$a = "abc"', "#a == 2"), b"abc abc"),
    # hex strings (test_hex_strings)
    (R("$a = { 31 32 [-] 38 39 }"), T1024 + b"1234567890 123456789 12 89"),
    (R("$a = { 31 32 [-] 33 34 [-] 38 39 }"), T1024 + b"1234567890 12 34 xx 89 12 89"),
    (R("$a = { 31 32 [250-300] 38 39 }"), b"12" + b"x" * 260 + b"89" + b"12" + b"y" * 100 + b"89"),
    (R("$a = { 31 32 ?? 34 [2-4] 38 (39|41) }"), T1024 + b"1234567890 1234xx89 1234xxxx8A"),
    (R("$a = { 64 01 00 00 60 01 }"), T1024 + bytes.fromhex("640100006001") * 3),
    (R("$a = { 6? 01 00 00 60 0? }"), T1024 + bytes.fromhex("640100006001") * 3),
    (R("$a = { 64 ~01 00 00 ~60 01 }"), bytes.fromhex("640200007001") + bytes.fromhex("640100006001")),
    (R("$a = { 64 ~0? 00 00 ~?0 01 }"), bytes.fromhex("641200006101") + bytes.fromhex("640100006001")),
    (R("$a = { AA AB AC AD [0-3] BB [2-3] CC }"), bytes.fromhex("AAABACADBB0000BBCC")),   # quirk 1 (row a8)
    (R("$a = { AA AB AC AD [0-3] BB [2-3] (CC|CC) }"), bytes.fromhex("AAABACADBB0000BBCC")),
    (R("$a = { 31 32 [0-2] 34 [0-2] 34 35 36 37 }"), b"12x4y4567 1244567 12xx4xx4567"),
    (R("$a = { 31 32 33 34 [1-2] 35 [1-2] 36 37 38 39 }"), b"1234x5x6789 1234xx5xx6789 1234x56789"),
    (R("$a = { 31 [1-3] 32 33 34 35 }"), b"1x2345 1xxx2345 12345 1xxxx2345"),
    # regular expressions (test_re)
    (R("$a = /ab[cd]+e{2,3}f/"), b"abcdcdeef abceeef abdeeeef xabceef"),
    (R("$a = /\\bfoo[a-z]*bar\\b/"), b"foobar fooxbar xfoobar foobarx foo bar fooooobar"),
    (R("$a = /a.{1,16}b/"), b"a123b45b aXb ab a\nb"),
    (R("$a = /a.*?b/ $b = /a.*b/", "$a and $b"), b"a123b45b aXb ab a\nb"),
    (R("$a = /abc/ nocase wide ascii"), b"ABC a\0B\0c\0 abc"),
    (R("$a = /^abc/ $b = /xyz$/"), b"abc abc xyz xyz"),
    (R("$a = /(abc|xyz|foo)+d/"), b"abcd xyzfood abcabcd food"),
    (R("$a = /a{2,5}?b/ $b = /a{2,5}b/"), b"aaaaaab aab ab aaab"),
    (R("$a = /[0-9a-f]{4,8}-w+\\w/"), b"12ab-www_ deadbeef-wX 123-ww"),
    (R("$a = /\\d+\\s\\D\\S\\W/"), b"123 a!. 4\tbc- 5  x"),
    (R("$a = /foo.{300,}bar/s"), b"foo" + b"z" * 310 + b"bar" + b"q" * 5 + b"bar"),   # greedy, !a == 324 vs 316
    (R("$a = /foo.{300,}?bar/s"), b"foo" + b"z" * 310 + b"bar" + b"q" * 5 + b"bar"),
    (R("$a = /abc[^d-f]x/ nocase"), b"abcGx abcdx ABCzX"),
    (R("$a = /\\Babc\\B/"), b"xabcx abc xabc abcx"),
    (R("$a = /a\\x00b\\x00c/ $b = /abc/ wide"), b"a\0b\0c\0"),
    (R("$a = /(a|ab)(c|bcd)(d*)/"), b"abcd abcdd acd"),
    (R("$a = /a*b/ $b = /(a*)*b/"), b"aaab b ab"),
    (R("$a = /abc/ fullword $b = /abc/ wide fullword"), b"abc xabc abc a\0b\0c\0 x\0a\0b\0c\0"),
    # work-area tiers of the GPU verifiers: > 96 live positions in the fast hex VM, > 64 live fibers in the fiber VM
    (R("$a = { 31 32 33 34 [0-180] 35 ?? [0-180] 36 37 38 39 }"), b"1234" + b"5" * 170 + b"x" * 20 + b"6789" + b"5z" * 40 + b"6789"),
    (R("$a = /abcd.{0,150}e.{0,150}fgh/s"), b"abcd" + b"e" * 140 + b"zz" + b"fgh" + b"e" * 30 + b"fgh"),
    (R("$a = /abcd(x|y|z)+.{0,90}q/"), b"abcd" + b"xyz" * 40 + b"q" + b"x" * 30 + b"q"),
    # backward-mode position lists keep duplicates (re.c:2323-2333): 52k live positions, far beyond the 4096-byte window
    (R("$a = { 61 [0-40] ?? [0-40] ?? [0-30] 62 63 64 65 }"), b"x" * 50 + b"a" * 200 + b"bcde" + b"y" * 20),
    # scan.c:1021-1023 for CHAINED strings: in fast mode, once a tail has completed the first chain, matches[head] is
    # non-empty and later candidates of the head string are not verified -- so the one that would exhaust the fiber
    # pool (ERROR_TOO_MANY_RE_FIBERS) never runs: rc 46 normally, rc 0 with one match in fast mode.  The other way
    # round (the exploding candidate first) both modes fail.
    (R("$s = /wxyz(a1(.??)+?c)??mn.{0,300}?QWERTY/", "$s"),
     b"...wxyzmn...." + b"-" * 50 + b"QWERTY" + b"." * 40 + b"wxyza1" + b"q" * 60 + b"QWERTY"),
    (R("$s = /wxyz(a1(.??)+?c)??mn.{0,300}?QWERTY/", "$s"),
     b"wxyza1" + b"q" * 60 + b"...wxyzmn...." + b"-" * 50 + b"QWERTY" + b"." * 40),
    # the same (string, offset) reported by two candidates (both "abcd" reach back to the same "x": scan.c:311-326 keeps
    # the first) in a scan that also completes a chain: the engine copies the plain strings' records to the host while
    # the chain replay is still running and must notice that the dedup moved the chain heads' records forward
    (R("$a = /x.{0,10}abcd/ $b = { 31 32 33 34 35 [250-300] 36 37 38 39 30 }", "$a and $b"),
     b"..x..abcd.abcd.." + b"12345" + b"-" * 260 + b"67890" + b"..x.abcdabcd" + b"12345" + b"+" * 290 + b"67890" + b"x" * 30),
]
