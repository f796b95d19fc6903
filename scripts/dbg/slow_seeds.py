import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 2 and sys.argv[1] == "one":
    import numpy as np, fuzzgen, yrh
    seed = int(sys.argv[2])
    R, B = yrh.ref(), yrh.b200()
    src, insts = fuzzgen.gen_rules(seed, 50)
    good = []
    for s in src:
        try: R.compile(s); good.append(s)
        except ValueError: pass
    rb = B.compile("".join(good))
    data = fuzzgen.gen_data(seed, insts, (1 << 16) + 17 * (seed % 100))
    rb.scan(data[:1000])
    t = time.time(); a = rb.scan(data, cap=1 << 20); print("seed", seed, "rc", a.rc, len(a.matches), "matches %.3f s" % (time.time() - t), flush=True)
    if len(sys.argv) > 3:
        small = data[: 1 << 14]
        t = time.time(); a = rb.scan_blocks(small, 1024, 256, cap=1 << 20); print("blocks %.3f s" % (time.time() - t), flush=True)
    sys.exit(0)
for seed in sys.argv[1:]:
    try:
        p = subprocess.run([sys.executable, __file__, "one"] + seed.split(":"), capture_output=True, text=True, timeout=90,
                           env=dict(os.environ, YRB_TIMING="1", YRB_TRACE="1"))
        print("==", seed, "rc", p.returncode); print(p.stdout[-800:]); print(p.stderr[-1800:])
    except subprocess.TimeoutExpired as e:
        print("==", seed, "TIMEOUT"); print((e.stdout or b"")[-800:]); print((e.stderr or b"")[-1800:])
