#!/bin/bash
# GPU session (not a test): the N=2 launch exactly as the round-end driver does it.
mkdir -p gpurun_out
export GZPB_BENCH_WATCHDOG=60
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --cpu-sample-mb 128 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?" >> gpurun_out/bench_n2.err
tail -n 30 gpurun_out/bench_n2.err | cut -c1-300; cat gpurun_out/bench_n2.json
