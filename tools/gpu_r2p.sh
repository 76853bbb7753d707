#!/bin/bash
# round 2, 8-GPU visit (final code, stage fusion on): weak-scaling bench with the strong-scaling and C5b sub-records, NCCL parity
TAG=${1:-r02p}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29522 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_bench8.json 2> $O/${TAG}_bench8.err; tail -1 $O/${TAG}_bench8.json | cut -c1-300; tail -2 $O/${TAG}_bench8.err
timeout 300 $TR --master-port 29521 tools/multigpu_check.py > $O/${TAG}_multigpu_check_8gpu.txt 2>&1; echo "multigpu_check exit $?"; grep -c " ok$" $O/${TAG}_multigpu_check_8gpu.txt; grep "FAIL\|PASSED\|Error" $O/${TAG}_multigpu_check_8gpu.txt | head
