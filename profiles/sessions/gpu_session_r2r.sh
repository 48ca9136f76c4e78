#!/bin/bash
# Round-2 session r (not a test), 2 GPUs: the final tree as the driver will run it — smoke(), GPU suite, bench N=1 and N=2 (both arms).
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_smoke.log
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "rc=$?" >> gpurun_out/r2r_bench_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference_n1.json 2> gpurun_out/r2r_bench_reference_n1.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "rc=$?" >> gpurun_out/r2r_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
    > gpurun_out/r2r_bench_reference_n2.json 2> gpurun_out/r2r_bench_reference_n2.err; echo "rc=$?" >> gpurun_out/r2r_bench_reference_n2.err
cat gpurun_out/r2r_smoke.log; tail -4 gpurun_out/pytest_gpu.log
for c in n1 n2; do python -c "
import json
for ln in open('gpurun_out/r2r_bench_$c.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'same', d['e2e'].get('identical_to_one_gpu_stream'), 'cpu', round(d['cpu_baseline']['value'],3), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'], d['clocks'])"; done
cut -c1-200 gpurun_out/r2r_bench_reference_n1.json; grep "^{" gpurun_out/r2r_bench_reference_n2.json | cut -c1-200
for f in gpurun_out/r2r_*.err; do tail -n 1 $f; done
