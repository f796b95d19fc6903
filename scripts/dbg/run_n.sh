YRB_TIMING=1 timeout 600 python bench.py --size-gib 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r02n_timing.json 2> gpurun_out/r02n_timing.err; grep "resolve" gpurun_out/r02n_timing.err | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_resolve_kernel -s 2 -c 1 -f -o gpurun_out/r02n_chain \
  python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/r02n_ncu_chain.log 2>&1; echo "ncu rc=$?"
