#!/bin/bash
# One GPU-box visit: parity tests, default bench, reference arm, ncu launch list, ncu --set full of the hot kernels.
# usage: tools/gpu_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 12 --launch-count 3 \
    -o $O/${TAG}_sweep -f python bench.py --n 256 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_qderiv3|k_rk_combine" --launch-skip 4 --launch-count 2 \
    -o $O/${TAG}_aux -f python bench.py --n 256 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_aux.log 2>&1
ls -la $O
