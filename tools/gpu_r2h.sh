#!/bin/bash
# 2 GPUs: NCCL parity check incl. compact schemes across ranks (line exchanges over NCCL), multi-rank drop-in tests
TAG=${1:-r02h}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/multigpu_check.py > $O/${TAG}_multigpu_check_2gpu.txt 2>&1; echo "multigpu_check exit $?"; grep -c " ok$" $O/${TAG}_multigpu_check_2gpu.txt; grep "FAIL\|compact\|PASSED\|Error" $O/${TAG}_multigpu_check_2gpu.txt | head -12
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "multirank" > $O/${TAG}_pytest_dropin_mp.log 2>&1; echo "dropin pytest exit $?"; tail -3 $O/${TAG}_pytest_dropin_mp.log
