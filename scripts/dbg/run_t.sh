timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded" 2>&1 | tail -2
bash scripts/r02_scale.sh 2 r02t
