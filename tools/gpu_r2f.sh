#!/bin/bash
# round 2: kernel-variant A/B (Q-derivative kernel stores / tiles, x-sweep stores), drop-in executable as a performance path
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
: > $O/${TAG}_variants.txt
for v in A B C D E F A B; do
  HYPAR_B200_LIB=$PWD/hypar_b200/csrc/variants/lib$v.so timeout 200 python bench.py --n 512 --steps 4 --warmup 2 --no-cpu --no-e2e --no-sub 2>/dev/null | tail -1 | \
    python -c "import sys,json; l=json.loads(sys.stdin.read()); s=l['roofline']['share_of_step']; ms=l['ms_per_step']; print('variant $v', round(l['value'],1), 'ms/step', round(ms,2), {k: round(v*ms,2) for k,v in s.items()}, l['clocks']['sm_mhz'])" >> $O/${TAG}_variants.txt
done
cat $O/${TAG}_variants.txt
timeout 900 python tools/dropin_bench.py --n 128 256 --steps 20 --out $O/${TAG}_dropin.json > $O/${TAG}_dropin.log 2>&1; tail -3 $O/${TAG}_dropin.log
