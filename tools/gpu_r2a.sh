#!/bin/bash
# round 2, first GPU visit (1 GPU): the library's distributed step over the in-process transport, default bench with
# sub-records, the reference's own CUDA path as second baseline, compute-sanitizer on the small cases.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_decomposed.py tests/test_gpu_diagnostics.py tests/test_parallel_io.py -m gpu -x -q > $O/${TAG}_pytest_decomposed.log 2>&1
echo "decomposed pytest exit $?"; tail -5 $O/${TAG}_pytest_decomposed.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_decomposed.py > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-600; tail -3 $O/${TAG}_bench.err
timeout 900 python tools/refgpu_bench.py --n 64 128 --steps 10 --out $O/${TAG}_refgpu.json > $O/${TAG}_refgpu.log 2>&1; tail -4 $O/${TAG}_refgpu.log
