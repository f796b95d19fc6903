#!/bin/bash
# The last GPU seconds of round 2 (the per-ruleset key split, commit "Discovery key: ..."): smoke, the new
# long-chain tests, then as much of the known-answer / engine parity tests as fits.  Per-test lines, so that
# a run cut off by the budget still says how far it got.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( python -u -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?" ) > gpurun_out/r02_last_check.log 2>&1
python -u -m pytest -p no:cacheprovider -m gpu -v \
  tests/test_zz_gpu_late_round2.py \
  "tests/test_gpu_parity.py::test_known_answer_cases" \
  "tests/test_gpu_parity.py::test_known_answer_cases_fast_mode" \
  "tests/test_gpu_parity.py::test_known_answer_cases_blocks" \
  tests/test_gpu_engine.py tests/test_gpu_batch.py >> gpurun_out/r02_last_check.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_last_check.log
