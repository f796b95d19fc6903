import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    import numpy as np, fuzzgen, yrh
    mode = sys.argv[1]
    B = yrh.b200()
    seed = 1002
    src, insts = fuzzgen.gen_rules(seed, 50)
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    r30 = [s for s in src if s.startswith("rule r30 ")][0]
    if mode == "min":
        rb = B.compile('rule r { strings: $s2 = /^zzyc.((\\b)*?)zbz/ condition: $s2 }')
        for d in (b"zzycazbz", b"zzyc.zbz"):
            t = time.time(); a = rb.scan(d); print(d, a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    elif mode == "r30":
        rb = B.compile(r30)
        t = time.time(); a = rb.scan(data); print("r30", a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    elif mode == "r30nore":
        import re
        rb = B.compile(re.sub(r"\$s2 = /[^\n]*\n", "", r30))
        t = time.time(); a = rb.scan(data); print("r30 without s2", a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    elif mode == "s2only":
        rb = B.compile('rule r { strings: $s2 = /^zzyc.((\\b)*?)zbz/ condition: $s2 }')
        t = time.time(); a = rb.scan(data); print("s2 only, full data", a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    elif mode == "others":
        good = []
        R = yrh.ref()
        for s in src:
            if s.startswith("rule r30 "): continue
            try: R.compile(s); good.append(s)
            except ValueError: pass
        rb = B.compile("".join(good))
        t = time.time(); a = rb.scan(data); print("others", a.rc, len(a.matches), "%.3f s" % (time.time() - t), flush=True)
    sys.exit(0)
for mode in ("s2only",):
    try:
        p = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True, timeout=25,
                           env=dict(os.environ, YRB_TIMING="1", YRB_DEBUG="1", YRB_TRACE="1"))
        print("==", mode, "rc", p.returncode); print(p.stdout[-1500:]); print(p.stderr[-1500:])
    except subprocess.TimeoutExpired as e:
        print("==", mode, "TIMEOUT"); print((e.stdout or b"")[-1500:]); print((e.stderr or b"")[-2500:])
