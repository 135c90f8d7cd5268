#!/bin/bash
# One round-end style pass on the GPU box: smoke, GPU parity tests, both bench arms, and the ncu
# launch list of the bench command.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
python tools/bench_ls.py --cpu-rows 2000000 | tee gpurun_out/bench_ls.json | cut -c1-600
if [ "$1" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 \
  > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
fi
