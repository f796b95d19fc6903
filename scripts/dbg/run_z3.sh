python - <<PY
import os, sys, time, subprocess, re
d = "/dev/shm/yrb_cli100k"
if not os.path.exists(d + "/rules.yarc"):
    subprocess.run([sys.executable, "scripts/cli_feeder.py", "100000", "600", "16", d], env=dict(os.environ, CLI_FEEDER_ONLY_PLAIN="1"), capture_output=True)
for p_, env in (("1", {}), ("1", {"YR_B200_RULE_PREFILTER": "0"}), ("16", {})):
    t = time.time()
    p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", p_, d + "/rules.yarc", d + "/files"], capture_output=True, text=True, env=dict(os.environ, YRB_TIMING="1", **env))
    dt = time.time() - t
    lines = [l for l in p.stderr.splitlines() if "block" in l]
    def col(name): return sorted(float(re.search(name + r" ([0-9.]+)", l).group(1)) for l in lines)
    e, g = col("engine"), col("this thread")
    q = lambda a, f: a[int(len(a) * f)]
    print("threads", p_, env, "total %.2f s" % dt, "files", len(lines), "engine p50 %.2f sum %.2f s | gap p50 %.2f p90 %.2f sum %.2f s" % (q(e,.5), sum(e)/1e3, q(g,.5), q(g,.9), sum(g)/1e3))
t = time.time(); subprocess.run(["oracle/_ref/yara_ref", "-C", "-p", "16", d + "/rules.yarc", d + "/files/blob00001.bin"], capture_output=True); print("reference CLI, one file: %.2f s" % (time.time() - t))
t = time.time(); subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", "16", d + "/rules.yarc", d + "/files/blob00001.bin"], capture_output=True); print("drop-in CLI, one file: %.2f s" % (time.time() - t))
PY
