#!/usr/bin/env python3
"""Where the time of a rules upload goes in a NEW process: CUDA context, flatten, yrb_rules_create
(image copy + filter tables from the registry / the disk cache / built), scanner.  Run it twice: the second process
finds the filter tables on disk.   python scripts/upload_timing.py [nstrings=20000] [yarc=/tmp/yrb_upload.yarc]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.time()
from yara_b200 import synth
from yara_b200.api import Engine, LibYara, DeviceBuffer, filter_cache_stats
t_import = time.time() - t0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
yarc = sys.argv[2] if len(sys.argv) > 2 else "/tmp/yrb_upload_%d.yarc" % n
ly = LibYara()
t = time.time()
if os.path.exists(yarc):
    rules, how = ly.load(yarc), "yr_rules_load"
else:
    rules, how = ly.compile(synth.ruleset(3, n).source), "compiled"
    ly.save(rules, yarc)
t_rules = time.time() - t
t = time.time(); b = DeviceBuffer(0, 1 << 20); t_ctx = time.time() - t      # first CUDA call of the process: context creation
t = time.time(); image = ly.flatten(rules); t_flat = time.time() - t
t = time.time(); eng = Engine(image, 0); t_up = time.time() - t              # yrb_rules_create + yrb_scanner_create
t = time.time(); eng2 = Engine(image, 0); t_up2 = time.time() - t            # same image again: registry
print(json.dumps({"strings": n, "rules": how, "rules_s": round(t_rules, 3), "cuda_context_s": round(t_ctx, 3), "flatten_s": round(t_flat, 3),
                  "rules_create_plus_scanner_s": round(t_up, 3), "second_upload_same_process_s": round(t_up2, 3),
                  "image_bytes": len(image), "filter_tables_from": filter_cache_stats(), "python_imports_s": round(t_import, 3)}))
