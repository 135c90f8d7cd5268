#!/bin/bash
# K3 A^T t: per-segment lane groups (MRB_LS_TMUL=seg) vs flat 32-entry windows (default)
for m in seg flat; do
  echo "== $m"
  MRB_LS_TMUL=$m python tools/bench_ls.py --cpu-rows 1000 | tee gpurun_out/bench_ls_$m.json | cut -c1-520
done
MRB_LS_TMUL=flat timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_ls_flat.csv \
  python tools/bench_ls.py --cpu-rows 1000 --no-warmup > gpurun_out/ncu_ls_flat.log 2>&1; echo "exit $?"
python tools/launch_summary.py gpurun_out/launches_ls_flat.csv | head -12
