#!/bin/bash
# sweep the filter-kernel variants (YRB_FILTER_VARIANT) on one GPU; prints value / ms / scan-kernel GB/s
for v in ${VARIANTS:-0 2 4}; do
  YRB_FILTER_VARIANT=$v python bench.py --size-gib ${SIZE_GIB:-8} --steps 5 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', 'value %.0f'%d['value'], 'ms %.3f'%d['ms_per_step'], 'scan GB/s %.0f'%d['roofline']['achieved'], d['config']['matches_rank0'])"
done
