#!/bin/bash
# round 2, 8-GPU visit: NCCL parity at 8 ranks (2x2x2), multi-rank drop-in executable, weak-scaling bench with the strong-scaling
# and C5b sub-records (overlapped schedule), the serial schedule for comparison, host-link probe
TAG=${1:-r02e}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 tools/multigpu_check.py > $O/${TAG}_multigpu_check_8gpu.txt 2>&1; echo "multigpu_check exit $?"; grep -c " ok$" $O/${TAG}_multigpu_check_8gpu.txt; grep -c "FAIL" $O/${TAG}_multigpu_check_8gpu.txt; tail -2 $O/${TAG}_multigpu_check_8gpu.txt
timeout 300 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "multirank or ensemble" > $O/${TAG}_pytest_dropin_mp.log 2>&1; echo "dropin pytest exit $?"; tail -3 $O/${TAG}_pytest_dropin_mp.log
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > $O/${TAG}_bench8.json 2> $O/${TAG}_bench8.err; tail -1 $O/${TAG}_bench8.json | cut -c1-300; tail -3 $O/${TAG}_bench8.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --serial-halo --no-e2e --no-cpu > $O/${TAG}_bench8_serial.json 2> $O/${TAG}_bench8_serial.err; tail -1 $O/${TAG}_bench8_serial.json | cut -c1-200
timeout 200 $TR --master-port 29524 tools/h2d_probe.py > $O/${TAG}_h2d_probe.json 2> $O/${TAG}_h2d_probe.err; tail -1 $O/${TAG}_h2d_probe.json
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1; lscpu | head -20 > $O/${TAG}_lscpu.txt
