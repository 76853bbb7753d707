#!/bin/bash
# round 2, 1-GPU visit: the owed store wait at the top of the next step (variant library) against after its P1 (default), same box
TAG=${1:-r02w}
O=gpurun_out
mkdir -p $O
HYPAR_B200_LIB=$PWD/hypar_b200/csrc/variants/libpendtop.so timeout 300 python -m pytest tests/test_gpu_stage_fusion.py tests/test_gpu_tma.py -m gpu -q -n 4 > $O/${TAG}_pytest_pendtop.log 2>&1; echo "pytest exit $?"; tail -2 $O/${TAG}_pytest_pendtop.log
: > $O/${TAG}_variants.txt
for v in default pendtop default pendtop; do
  L=$PWD/hypar_b200/libhypar_b200.so; [ $v = pendtop ] && L=$PWD/hypar_b200/csrc/variants/libpendtop.so
  HYPAR_B200_LIB=$L timeout 200 python bench.py --n 512 --steps 6 --warmup 3 --no-cpu --no-e2e --no-sub 2>/dev/null | tail -1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.read()); s=l['roofline']['share_of_step']; ms=l['ms_per_step']; print('variant $v', round(l['value'],1), 'ms/step', round(ms,2), {k: round(v*ms,2) for k,v in s.items()}, l['clocks']['sm_mhz'])" >> $O/${TAG}_variants.txt
done
cat $O/${TAG}_variants.txt
