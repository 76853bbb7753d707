#!/bin/bash
# One GPU-box visit: parity tests, default bench, reference arm, ncu launch list, ncu --set full of the hot kernels.
# usage: tools/gpu_round.sh <tag> [skip-tests]     (outputs under gpurun_out/<tag>_*)
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest_gpu.log
  tail -3 $O/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json | cut -c1-300
# launch list of the bench command itself (device-resident loop; the e2e / cpu legs add no kernels of other kinds)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
# the dominant kernel at the bench's own size: one launch per direction of the second stage
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 --launch-count 3 \
    -o $O/${TAG}_sweep512 -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_qderiv_int|k_rk_combine" --launch-skip 2 --launch-count 2 \
    -o $O/${TAG}_aux -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_aux.log 2>&1
ls -la $O | tail -12
