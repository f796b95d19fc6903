python - <<PY
import os, sys, time, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from yara_b200 import synth
from yara_b200.api import LibYara
d = "/dev/shm/yrb_cli_small"; os.makedirs(d + "/files", exist_ok=True)
rs = synth.ruleset(4, 100000)
import random, math
rng = random.Random(3)
for k in range(24):
    s = int(math.exp(rng.uniform(math.log(65536), math.log(8 << 20))))
    synth.corpus(rs, s, 65536, seed=1000 + k, plant_seed=k, tail_plant=False).tofile("%s/files/b%03d.bin" % (d, k))
ly = LibYara(); ly.save(ly.compile(rs.source), d + "/rules.yarc")
for p_, env in (("1", {}), ("16", {}), ("16", {"YR_B200_COALESCE": "16"})):
    t = time.time()
    p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", p_, d + "/rules.yarc", d + "/files"], capture_output=True, text=True, env=dict(os.environ, YRB_TIMING="1", **env))
    dt = time.time() - t
    lines = [l for l in p.stderr.splitlines() if "block" in l]
    print("threads", p_, env, "total %.2f s" % dt, "rc", p.returncode, len(lines), "timing lines")
    for l in lines[:3] + lines[-6:]: print("   ", l[:260])
t = time.time(); p = subprocess.run(["oracle/_ref/yara_ref", "-C", "-p", "1", d + "/rules.yarc", d + "/files"], capture_output=True, text=True); print("ref -p 1 total %.2f s" % (time.time() - t))
t = time.time(); p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", "1", d + "/rules.yarc", d + "/files/b000.bin"], capture_output=True, text=True); print("drop-in, ONE file: total %.2f s" % (time.time() - t))
PY
