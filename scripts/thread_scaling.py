#!/usr/bin/env python3
"""Per-file scans through the drop-in with 1..16 scanner threads (the shape of `yara -p N`)."""
import subprocess, sys
p = subprocess.run([sys.executable, "scripts/file_batch.py", "20000", "300", "1,2,4,8,16", "--no-batch"], capture_output=True, text=True)
print("".join(l + "\n" for l in p.stdout.splitlines() if l.startswith(("drop-in", "reference"))), end="")
if p.returncode:
    print(p.stderr[-500:])
