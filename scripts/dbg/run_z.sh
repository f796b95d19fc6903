python - <<PY
import os, sys, time, subprocess
d = "/dev/shm/yrb_cli_small"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from yara_b200 import synth
from yara_b200.api import LibYara
os.makedirs(d + "/files", exist_ok=True)
rs = synth.ruleset(4, 100000)
import random, math
rng = random.Random(3)
for k in range(24):
    s = int(math.exp(rng.uniform(math.log(65536), math.log(8 << 20))))
    synth.corpus(rs, s, 65536, seed=1000 + k, plant_seed=k, tail_plant=False).tofile("%s/files/b%03d.bin" % (d, k))
ly = LibYara(); ly.save(ly.compile(rs.source), d + "/rules.yarc")
for env in ({}, {"YRB_POOLED_ALLOC": "0"}):
    t = time.time()
    p = subprocess.run(["oracle/_ref/yara_b200cli", "-C", "-p", "16", d + "/rules.yarc", d + "/files"], capture_output=True, text=True, env=dict(os.environ, YRB_TIMING="1", **env))
    dt = time.time() - t
    lines = [l for l in p.stderr.splitlines() if "block" in l]
    import re
    ms = sorted(float(re.search(r"engine ([0-9.]+) ms", l).group(1)) for l in lines)
    print("threads 16", env, "total %.2f s" % dt, "rc", p.returncode, "engine ms per file: median %.1f max %.1f" % (ms[len(ms)//2], ms[-1]))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -x -q -k "known_answer or concurrent or batch" 2>&1 | tail -2
timeout 1200 python scripts/cli_feeder.py 100000 600 16 /dev/shm/yrb_cli100k > gpurun_out/r02z_cli_feeder_100k.jsonl 2> gpurun_out/r02z_cli_feeder_100k.err; echo "cli rc=$?"; cut -c1-200 gpurun_out/r02z_cli_feeder_100k.jsonl
timeout 600 python bench.py --no-cpu --no-parity > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; python - <<PY
import json
for l in open("gpurun_out/r02z_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("bench", round(d["value"],1), round(d["ms_per_step"],3), d["config"]["ms_breakdown"], d["e2e"]["value"])
PY
