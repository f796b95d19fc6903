python - <<PY
import os, sys, time, subprocess, re
d = "/dev/shm/yrb_cli100k"
if not os.path.exists(d + "/rules.yarc"):
    subprocess.run([sys.executable, "scripts/cli_feeder.py", "100000", "600", "16", d], env=dict(os.environ, CLI_FEEDER_ONLY_PLAIN="1"), capture_output=True)
for p_ in ("16", "4", "1"):
    t = time.time()
    p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", p_, d + "/rules.yarc", d + "/files"], capture_output=True, text=True, env=dict(os.environ, YRB_TIMING="1"))
    dt = time.time() - t
    lines = [l for l in p.stderr.splitlines() if "block" in l]
    def col(name): return sorted(float(re.search(name + r" ([0-9.]+)", l).group(1)) for l in lines)
    e, sc, ve, rs_, ma = col("engine"), col(r"scan\+h2d"), col("verify"), col("resolve"), col("materialise")
    q = lambda a, f: a[int(len(a) * f)]
    print("threads", p_, "total %.2f s" % dt, "files", len(lines), "engine ms p50 %.2f p90 %.2f p99 %.2f sum %.1f s | p50 scan+h2d %.2f verify %.2f resolve %.2f materialise %.3f" % (q(e,.5), q(e,.9), q(e,.99), sum(e)/1e3, q(sc,.5), q(ve,.5), q(rs_,.5), q(ma,.5)))
PY
