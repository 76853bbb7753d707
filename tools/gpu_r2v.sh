#!/bin/bash
# round 2, 1-GPU visit: k_qderiv_int with 64 x 8 tiles (variant library) against the 32 x 16 default on the same box
TAG=${1:-r02v}
O=gpurun_out
mkdir -p $O
HYPAR_B200_LIB=$PWD/hypar_b200/csrc/variants/libq64x8.so timeout 300 python -m pytest tests/test_gpu_stage_fusion.py tests/test_gpu_tma.py -m gpu -q -n 4 > $O/${TAG}_pytest_q64x8.log 2>&1; echo "pytest exit $?"; tail -2 $O/${TAG}_pytest_q64x8.log
: > $O/${TAG}_variants.txt
for v in default q64x8 default q64x8; do
  L=$PWD/hypar_b200/libhypar_b200.so; [ $v = q64x8 ] && L=$PWD/hypar_b200/csrc/variants/libq64x8.so
  HYPAR_B200_LIB=$L timeout 200 python bench.py --n 512 --steps 6 --warmup 3 --no-cpu --no-e2e --no-sub 2>/dev/null | tail -1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.read()); s=l['roofline']['share_of_step']; ms=l['ms_per_step']; print('variant $v', round(l['value'],1), 'ms/step', round(ms,2), {k: round(v*ms,2) for k,v in s.items()}, l['clocks']['sm_mhz'])" >> $O/${TAG}_variants.txt
done
cat $O/${TAG}_variants.txt
