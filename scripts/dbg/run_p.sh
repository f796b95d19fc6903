timeout 900 python -m pytest tests/test_reference_suite.py tests/test_fuzz.py tests/test_gpu_parity.py -m gpu -x -q -k "reference_program or blocks or known_answer" > gpurun_out/r02p_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02p_gputests.log
cd oracle/_ref; export TOP_SRCDIR=$PWD/top
( time ./test-rules_b200 > /dev/null 2>&1 ) 2>&1 | grep real
( time YR_B200_BLOCK_BATCH=0 ./test-rules_b200 > /dev/null 2>&1 ) 2>&1 | grep real
( time ./test-rules_ref > /dev/null 2>&1 ) 2>&1 | grep real
cd ../..
timeout 1500 python scripts/block_batch.py 2000 16 4096 > gpurun_out/r02p_block_batch.jsonl 2> gpurun_out/r02p_block_batch.err; echo "rc=$?"; cat gpurun_out/r02p_block_batch.jsonl; tail -3 gpurun_out/r02p_block_batch.err
timeout 1500 python scripts/block_batch.py 2000 8 1024 >> gpurun_out/r02p_block_batch.jsonl 2>> gpurun_out/r02p_block_batch.err; echo "rc=$?"; tail -3 gpurun_out/r02p_block_batch.jsonl
