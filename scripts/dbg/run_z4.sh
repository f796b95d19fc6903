python - <<PY
import os, sys, time, subprocess, re
d = "/dev/shm/yrb_cli100k"
if not os.path.exists(d + "/rules.yarc"):
    subprocess.run([sys.executable, "scripts/cli_feeder.py", "100000", "600", "16", d], env=dict(os.environ, CLI_FEEDER_ONLY_PLAIN="1"), capture_output=True)
for rep in range(2):
  for p_ in ("1", "16"):
    t = time.time()
    p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", p_, d + "/rules.yarc", d + "/files/blob00001.bin"], capture_output=True, text=True, env=dict(os.environ, YRB_TIMING="1"))
    dt = time.time() - t
    print("one file, -p", p_, "total %.2f s |" % dt, " | ".join(l[11:140] for l in p.stderr.splitlines() if "rules of" in l or "block" in l))
t = time.time(); subprocess.run([sys.executable, "-c", "import ctypes; l=ctypes.CDLL('yara_b200/libyrb200.so'); print(l.yrb_device_count())"], capture_output=True); print("python + dlopen + device count: %.2f s" % (time.time() - t))
PY
