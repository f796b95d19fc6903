import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
# usage: seed_rule.py SEED RULENAME [keep-only-string e.g. s0]
if sys.argv[1] == "one":
    import re
    import numpy as np, fuzzgen, yrh
    seed, name = int(sys.argv[2]), sys.argv[3]
    R, B = yrh.ref(), yrh.b200()
    src, insts = fuzzgen.gen_rules(seed, 50)
    rule = [s for s in src if s.startswith("rule %s " % name)][0]
    if len(sys.argv) > 4:
        keep = sys.argv[4]
        lines = [l for l in rule.split("\n") if not l.strip().startswith("$") or l.strip().startswith("$" + keep + " ")]
        rule = "\n".join(lines)
        rule = re.sub(r"condition:.*?\}", "condition: any of them }", rule, flags=re.S)
    print(rule)
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    rr, rb = R.compile(rule), B.compile(rule)
    t = time.time(); a = rr.scan(data, cap=1 << 20); print("ref rc", a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    t = time.time(); b = rb.scan(data, cap=1 << 20); print("gpu rc", b.rc, len(b.matches), "%.3f s" % (time.time() - t), flush=True)
    sys.exit(0)
try:
    p = subprocess.run([sys.executable, __file__, "one"] + sys.argv[1:], capture_output=True, text=True, timeout=100,
                       env=dict(os.environ, YRB_TIMING="1", YRB_TRACE="1"))
    print(p.stdout[-1500:]); print(p.stderr[-2500:])
except subprocess.TimeoutExpired as e:
    print("TIMEOUT"); print((e.stdout or b"")[-1500:]); print((e.stderr or b"")[-2500:])
