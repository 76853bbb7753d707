#!/bin/bash
# round 2, 1-GPU visit: stage fusion (the last sweep of an RK stage writes the next stage solution) -- bit-identity tests,
# the sweeps' own tests, bench with and without it, sanitizers on the small cases (which now run the RKF kernel)
TAG=${1:-r02k}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stage_fusion.py tests/test_gpu_tma.py tests/test_gpu_roe_fused.py -m gpu -q -x > $O/${TAG}_pytest_fusion.log 2>&1; echo "pytest exit $?"; tail -15 $O/${TAG}_pytest_fusion.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu > $O/${TAG}_bench_fused.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_fused.json | cut -c1-300
HPB_STAGE_FUSION=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu > $O/${TAG}_bench_unfused.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_unfused.json | cut -c1-300
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py quick > $O/${TAG}_sanitizer_$tool.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|SANITIZE CASES OK|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.txt | tail -3
done
