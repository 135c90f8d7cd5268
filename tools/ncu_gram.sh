#!/bin/bash
# Fresh ncu --set full capture (with source) of the two k_gram launches of one C3 sweep.
# The same command runs once WITHOUT ncu first and must exit 0.
mkdir -p gpurun_out
CMD="python tools/profile_als.py 283228 53889 27753444 50 4 2"
$CMD > gpurun_out/profile_als_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/profile_als_plain.log; exit 1; }
tail -1 gpurun_out/profile_als_plain.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_gram --launch-skip 2 --launch-count 2 \
    -f -o gpurun_out/ncu_k_gram_C3_r02 $CMD > gpurun_out/ncu_k_gram.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_k_gram.log; ls -la gpurun_out/*.ncu-rep
