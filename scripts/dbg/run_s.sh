python scripts/sanitize_smoke.py 2>&1 | tail -6
timeout 900 python scripts/fuzz_sweep.py 1000 1140 420 "1002,1189,1254,1306" > gpurun_out/r02s_fuzz.log 2> gpurun_out/r02s_fuzz.err; echo "fuzz rc=$?"; tail -1 gpurun_out/r02s_fuzz.log
