#!/bin/bash
# round 2, 1-GPU visit: the deferred store wait in the accumulating sweeps -- sweep tests, bit-identity of the stage fusion,
# decomposed runs, racecheck / memcheck on the small cases, bench with and without stage fusion
TAG=${1:-r02m}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stage_fusion.py tests/test_gpu_tma.py tests/test_gpu_roe_fused.py tests/test_gpu_decomposed.py tests/test_gpu_parity.py -m gpu -q -n 4 > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 $O/${TAG}_pytest.log
for tool in racecheck memcheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py quick > $O/${TAG}_sanitizer_$tool.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|SANITIZE CASES OK|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.txt | tail -3
done
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_fused.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_fused.json | cut -c1-200
HPB_STAGE_FUSION=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_unfused.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_unfused.json | cut -c1-200
