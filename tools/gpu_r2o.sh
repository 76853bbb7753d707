#!/bin/bash
# round 2, 2-GPU visit: NCCL parity check (stage fusion on by default; compact schemes incl. the characteristic ones, hcweno5 and
# GLM-GEE across ranks), multi-rank drop-in tests, bench on 2 GPUs
TAG=${1:-r02o}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/multigpu_check.py > $O/${TAG}_multigpu_check_2gpu.txt 2>&1; echo "multigpu_check exit $?"; grep -c " ok$" $O/${TAG}_multigpu_check_2gpu.txt; grep "FAIL\|PASSED\|Error" $O/${TAG}_multigpu_check_2gpu.txt | head -12
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "multirank" > $O/${TAG}_pytest_dropin_mp.log 2>&1; echo "dropin pytest exit $?"; tail -3 $O/${TAG}_pytest_dropin_mp.log
timeout 600 $TR --master-port 29532 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > $O/${TAG}_bench2.json 2> $O/${TAG}_bench2.err; tail -1 $O/${TAG}_bench2.json | cut -c1-300; tail -3 $O/${TAG}_bench2.err
