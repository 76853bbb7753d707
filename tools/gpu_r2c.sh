#!/bin/bash
# round 2, 1-GPU visit: remaining tests, parity table, sanitizers, launch lists (ours and the reference's CUDA path),
# ncu --set full of the sweeps and the Q-derivative kernel
TAG=${1:-r02c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_roe_fused.py tests/test_gpu_tma.py tests/test_gpu_decomposed.py tests/test_gpu_ensemble.py -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/${TAG}_pytest.log
timeout 300 python tools/parity_table.py --out $O/${TAG}_parity_table.txt > /dev/null 2>$O/${TAG}_parity.err; tail -3 $O/${TAG}_parity_table.txt
for tool in memcheck racecheck synccheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py quick > $O/${TAG}_sanitizer_$tool.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|SANITIZE CASES OK|RACECHECK SUMMARY" $O/${TAG}_sanitizer_$tool.txt | tail -3
done
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_bench.log 2>&1
# the reference's own CUDA path: Mpoint-RK-stage/s at 64^3 .. 256^3, and its launch list (one RK4 step at 64^3)
timeout 900 python tools/refgpu_bench.py --n 128 256 --steps 6 --out $O/${TAG}_refgpu.json > $O/${TAG}_refgpu.log 2>&1; tail -3 $O/${TAG}_refgpu.log
timeout 600 python tools/refgpu_bench.py --n 128 256 --steps 6 --upwinding roe --tstype ssprk3 --out $O/${TAG}_refgpu_roe.json > $O/${TAG}_refgpu_roe.log 2>&1; tail -3 $O/${TAG}_refgpu_roe.log
python tools/refgpu_bench.py --n 64 --steps 2 --prepare /tmp/refgpu_dir
(cd /tmp/refgpu_dir && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OLDPWD/$O/${TAG}_refgpu_launches.csv $OLDPWD/oracle/_ref/hypar_ref_gpu > $OLDPWD/$O/${TAG}_refgpu_ncu.log 2>&1)
# ncu --set full: one launch per direction of the second stage (Rusanov C4), the Q-derivative kernel, the Roe sweeps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 --launch-count 3 \
    -o $O/${TAG}_sweep512 -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_qderiv_int|k_rk_combine" --launch-skip 2 --launch-count 2 \
    -o $O/${TAG}_aux -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_aux.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 3 --launch-count 3 \
    -o $O/${TAG}_sweep512_roe -f python bench.py --n 512 --steps 1 --warmup 1 --no-cpu --no-e2e --no-sub --workload c4roe > $O/${TAG}_ncu_sweep_roe.log 2>&1
ls -la $O | tail -20
