#!/bin/bash
# Multi-GPU pass on an N-GPU box (gpurun --gpus N): the real 2-process parity test, the headline
# bench under torchrun, config 4 split by query block, config 5 at a reduced scale.
# usage: tools/gpu_multi.sh N [c5_scale]
N=$1; S=${2:-0.02}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
if [ "$N" == "2" ]; then
  ( timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q ) 2>&1 | tail -5 | tee gpurun_out/pytest_sharded_n2.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
MRB_E2E_TIMING=1 timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
cut -c1-1200 gpurun_out/bench_n$N.json; grep -E "e2e N=|ShardedAls|Error|error" gpurun_out/bench_n$N.err | tail -8
timeout 600 $TR bench.py --gpus $N --config C4 --steps 3 --warmup 1 > gpurun_out/bench_C4_n$N.json 2> gpurun_out/bench_C4_n$N.err; echo "C4 N=$N exit $?"
cut -c1-500 gpurun_out/bench_C4_n$N.json; tail -3 gpurun_out/bench_C4_n$N.err
timeout 900 $TR bench.py --gpus $N --config C5 --scale $S --steps 2 --warmup 1 > gpurun_out/bench_C5_n$N.json 2> gpurun_out/bench_C5_n$N.err; echo "C5 N=$N exit $?"
cut -c1-1600 gpurun_out/bench_C5_n$N.json; tail -5 gpurun_out/bench_C5_n$N.err
