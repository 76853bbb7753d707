#!/bin/bash
# round 2, 2-GPU visit: NCCL transport inside the library -- parity over NCCL, both schedules, multi-rank drop-in executable,
# bench at 2 GPUs (overlapped / serial / without buffer registration)
TAG=${1:-r02b}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multigpu_check.py > $O/${TAG}_multigpu_check_2gpu.txt 2>&1; echo "multigpu_check exit $?"; tail -4 $O/${TAG}_multigpu_check_2gpu.txt
timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_roe_fused.py tests/test_gpu_tma.py tests/test_gpu_decomposed.py -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/${TAG}_pytest.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/${TAG}_bench2.json 2> $O/${TAG}_bench2.err; tail -1 $O/${TAG}_bench2.json | cut -c1-300; tail -3 $O/${TAG}_bench2.err
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --serial-halo --no-e2e --no-cpu --no-sub > $O/${TAG}_bench2_serial.json 2> $O/${TAG}_bench2_serial.err; tail -1 $O/${TAG}_bench2_serial.json | cut -c1-200
HPB_NCCL_REGISTER=0 timeout 400 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --no-sub > $O/${TAG}_bench2_noreg.json 2> $O/${TAG}_bench2_noreg.err; tail -1 $O/${TAG}_bench2_noreg.json | cut -c1-200
