#!/bin/bash
# ncu --set full captures of the dominant kernels of K3 (config 2), K6 (co-rating similarity) and
# K7 (data preparation) at full size; summaries are made afterwards with tools/ncu_summary.py.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:'k_csc_seg_native|k_csc_tmul_warp|k_csr_mul' -s 12 -c 4 -o gpurun_out/ncu_k3_C2 \
  python tools/bench_ls.py --cpu-rows 1000 > gpurun_out/ncu_k3.log 2>&1; echo "k3 exit $?"
timeout 400 $NCU -k regex:k_cosim -s 1 -c 1 -o gpurun_out/ncu_k6 \
  python tools/bench_cosim.py --queries 8192 > gpurun_out/ncu_k6.log 2>&1; echo "k6 exit $?"
timeout 400 $NCU -k regex:'k_count_alive|k_compact|k_rating_keys|k_radix_scatter' -s 10 -c 8 -o gpurun_out/ncu_k7 \
  python tools/bench_prep.py 120 2000 > gpurun_out/ncu_k7.log 2>&1; echo "k7 exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_ls_native.csv \
  python tools/bench_ls.py --cpu-rows 1000 > gpurun_out/ncu_ls_launches.log 2>&1; echo "ls launches exit $?"
ls -la gpurun_out/*.ncu-rep
