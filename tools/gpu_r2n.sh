#!/bin/bash
# round 2, 1-GPU visit: bench with and without stage fusion (same box, back to back), reference arm, final ncu --set full of the
# sweeps of the second stage and of the remaining stage kernels
TAG=${1:-r02n}
O=gpurun_out
mkdir -p $O
timeout 500 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_fused.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_fused.json | cut -c1-200
HPB_STAGE_FUSION=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_unfused.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_unfused.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 --launch-count 3 \
    -o $O/${TAG}_sweep512 -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_qderiv_int|k_rk_combine" --launch-skip 2 --launch-count 2 \
    -o $O/${TAG}_aux -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_aux.log 2>&1
ls -la $O | tail -6
