bash scripts/r02_evidence.sh r02u
timeout 1200 python scripts/cli_feeder.py 100000 600 16 /dev/shm/yrb_cli100k > gpurun_out/r02u_cli_feeder_100k.jsonl 2> gpurun_out/r02u_cli_feeder_100k.err; echo "cli rc=$?"; cut -c1-260 gpurun_out/r02u_cli_feeder_100k.jsonl
