#!/bin/bash
# K3 (config 2, native-order sparse LS) at full size: launch list + ncu --set full of its three operator kernels
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ls_native.csv \
  python tools/bench_ls.py --cpu-rows 1000 --no-warmup > gpurun_out/ncu_ls_launches.log 2>&1; echo "ls launches exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:'k_csc_seg_native|k_csr_mul|k_csc_fold' -s 3 -c 3 \
  -o gpurun_out/ncu_k3_C2 python tools/bench_ls.py --cpu-rows 1000 --no-warmup > gpurun_out/ncu_k3.log 2>&1; echo "k3 exit $?"
python tools/bench_ls.py --cpu-rows 2000000 | tee gpurun_out/bench_ls.json | cut -c1-600
python tools/bench_prep.py 120 2000 | tee gpurun_out/bench_prep3.json | cut -c1-400
timeout 200 python -m pytest tests/test_gpu_prep.py -x -q 2>&1 | tail -2
