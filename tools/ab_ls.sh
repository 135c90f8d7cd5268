#!/bin/bash
# K3 A^T t: per-segment lane groups (MRB_LS_TMUL=seg) vs flat 32-entry windows (default)
for m in seg flat; do
  echo "== $m"; MRB_LS_TMUL=$m timeout 300 python -m pytest tests/test_gpu_ls.py -x -q 2>&1 | tail -2
  MRB_LS_TMUL=$m python tools/bench_ls.py --cpu-rows 1000 | tee gpurun_out/bench_ls_$m.json | cut -c1-420
done
