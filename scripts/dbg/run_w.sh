df -h /dev/shm /tmp | cat
CLI_FEEDER_ONLY_PLAIN=1 timeout 1200 python scripts/cli_feeder.py 100000 600 16 /tmp/yrb_cli100k > gpurun_out/r02w_cli.jsonl 2> gpurun_out/r02w_cli.err; echo "cli rc=$?"; cut -c1-200 gpurun_out/r02w_cli.jsonl; tail -3 gpurun_out/r02w_cli.err | cut -c1-500
df -h /dev/shm /tmp | cat
