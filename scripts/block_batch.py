#!/usr/bin/env python3
"""multi-block scans through the drop-in (yr_scanner_scan_mem_blocks with a block iterator, what
process-memory scans use: scanner.c:467-500, proc/linux.c:309) with the small blocks batched (default) and one
engine pipeline per block ($YR_B200_BLOCK_BATCH=0), against the reference library on one host thread.
  python scripts/block_batch.py [nstrings=2000] [MiB=16] [block=4096]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, time, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import yrh
from yara_b200 import synth
n, mib, blk, which = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rs = synth.ruleset(3, n)
lib = yrh.ref() if which == "ref" else yrh.b200()
r = lib.compile(rs.source)
data = synth.corpus(rs, mib << 20, 65536)
r.scan_blocks(data[: 1 << 20], blk, 256, cap=1 << 20)          # warm-up: device scanner, CUDA context
t = time.time(); res = r.scan_blocks(data, blk, 256, cap=1 << 20); dt = time.time() - t
import hashlib
print(json.dumps({"arm": which, "seconds": round(dt, 3), "blocks": (len(data) + blk - 257) // (blk - 256), "MiB": mib, "block_bytes": blk,
                  "matches": len(res.matches), "rc": res.rc, "sha1": hashlib.sha1(res.matches.tobytes()).hexdigest()[:16]}))
"""
n = sys.argv[1] if len(sys.argv) > 1 else "2000"
mib = sys.argv[2] if len(sys.argv) > 2 else "16"
blk = sys.argv[3] if len(sys.argv) > 3 else "4096"
code = CODE % (ROOT, os.path.join(ROOT, "tests"))
outs = []
for which, env in (("ref", {}), ("b200", {}), ("b200", {"YR_B200_BLOCK_BATCH": "0"})):
    p = subprocess.run([sys.executable, "-c", code, n, mib, blk, which], capture_output=True, text=True, env=dict(os.environ, **env), timeout=3000)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and line, (p.stdout + p.stderr)[-2000:]
    d = json.loads(line[0]); d["env"] = env
    outs.append(d)
    print(json.dumps(d), flush=True)
assert len(set(d["sha1"] for d in outs)) == 1, "results differ"
