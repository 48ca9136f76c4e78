#!/bin/bash
# GPU session (not a test): new writer / ticket / full-size tests first, writer throughput, bench, then the rest of the GPU suite.
mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_writer.py tests/test_gpu_fullsize.py -m gpu -q --tb=short ) > gpurun_out/pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_new.log
timeout 70 python tests/perf_writer.py > gpurun_out/perf_writer.json 2> gpurun_out/perf_writer.err
echo "rc=$?" >> gpurun_out/perf_writer.err
timeout 100 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?" >> gpurun_out/bench.err
( time timeout 150 python -m pytest tests -m gpu -q --tb=short --ignore=tests/test_gpu_writer.py --ignore=tests/test_gpu_fullsize.py ) > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 40 gpurun_out/pytest_new.log; cat gpurun_out/perf_writer.json; tail -3 gpurun_out/perf_writer.err; cat gpurun_out/bench.json | cut -c1-700; tail -3 gpurun_out/bench.err; tail -n 12 gpurun_out/pytest_gpu.log
