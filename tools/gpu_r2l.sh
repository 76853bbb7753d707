#!/bin/bash
# round 2, 1-GPU visit: the whole GPU suite with stage fusion on by default, launch list of the bench command, ncu --set full
# of the sweeps of the second stage (x, y, fused z) and of the remaining stage kernels
TAG=${1:-r02l}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 --launch-count 3 \
    -o $O/${TAG}_sweep512 -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_sweep.log 2>&1
ls -la $O | tail -8
