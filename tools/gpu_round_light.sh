#!/bin/bash
# A GPU-box visit without the ncu --set full captures: parity tests, default bench, reference arm, ncu launch list.
# usage: tools/gpu_round_light.sh <tag>      (outputs under gpurun_out/<tag>_*)
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
ls -la $O | tail -8
