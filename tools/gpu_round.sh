#!/bin/bash
# One round-end style pass on the GPU box: smoke, GPU parity tests, both bench arms, the other
# configurations.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cut -c1-1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for c in C2 C4 prep a8; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 1 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
  cut -c1-700 gpurun_out/bench_$c.json; tail -3 gpurun_out/bench_$c.err
done
if [ "$1" == "ref" ]; then
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
fi
