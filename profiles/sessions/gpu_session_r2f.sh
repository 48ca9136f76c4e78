#!/bin/bash
# Round-2 session f (not a test), on a multi-GPU box (gpurun --gpus N): one ordered stream over N GPUs.
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/sessions/gpu_session_r2f.sh 2'
N=${1:-2}
mkdir -p gpurun_out
( nvidia-smi -L; nproc; free -g | head -2; nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name" ) > gpurun_out/r2k_host_n$N.txt 2>&1
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -k "multi or all_gpus or over_all" ) > gpurun_out/r2k_pytest_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest_n$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 \
    > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2k_bench_n$N.err
for feed in reserve write; do
GZPB_FULL_COPIES=${COPIES:-4000} timeout 300 python bench.py --full-stream --gpus $N --feed $feed > gpurun_out/r2k_fullstream_${feed}_n$N.json 2> gpurun_out/r2k_fullstream_${feed}_n$N.err; echo "rc=$?" >> gpurun_out/r2k_fullstream_${feed}_n$N.err
done
tail -4 gpurun_out/r2k_pytest_n$N.log
head -c 2500 gpurun_out/r2k_bench_n$N.json; echo
cat gpurun_out/r2k_fullstream_*_n$N.json
tail -n 5 gpurun_out/r2k_bench_n$N.err
tail -n 3 gpurun_out/r2k_fullstream_*_n$N.err
