#!/bin/bash
# A/B of the scheduler's ticket order (MRB_WORK_ORDER) on the full C3 workload + correctness of mode 1
mkdir -p gpurun_out
for rep in 1 2; do
  for m in 0 1; do
    echo -n "order $m: "; MRB_WORK_ORDER=$m python tools/profile_als.py 283228 53889 27753444 50 4 10 | tail -1
  done
done
MRB_WORK_ORDER=1 timeout 600 python -m pytest tests/test_gpu_als_gram.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
python tools/bench_prep.py 120 2000 | python -c "import json,sys; d=json.load(sys.stdin); print('medians kernel_ms', d['medians']['kernel_ms'], 'shrink', d['shrink']['kernel_ms'])"
