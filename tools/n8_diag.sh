TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
MRB_E2E_TIMING=1 timeout 200 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-parity > gpurun_out/diag_owned.json 2> gpurun_out/diag_owned.err; echo "owned exit $?"
grep -h "timed out" gpurun_out/diag_owned.err | cut -c1-400 | sort | uniq -c | head -12
MRB_FULL_INDEX=1 MRB_E2E_TIMING=1 timeout 200 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-parity > gpurun_out/diag_full.json 2> gpurun_out/diag_full.err; echo "full exit $?"
grep -h "timed out" gpurun_out/diag_full.err | cut -c1-400 | sort | uniq -c | head -12
cut -c1-300 gpurun_out/diag_full.json
