#!/bin/bash
# round 2, 1-GPU visit: viscosity law as a Halley root (HPB_MU_FAST) -- the whole GPU suite, the bench line exactly as the driver
# runs it (timed), the same with the exp/log viscosity law (variant library) on the same box
TAG=${1:-r02s}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
SECONDS=0; timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench wall ${SECONDS}s"; tail -1 $O/${TAG}_bench.json | cut -c1-250
HYPAR_B200_LIB=$PWD/hypar_b200/csrc/variants/libmuslow.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-sub > $O/${TAG}_bench_muslow.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_muslow.json | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-sub > $O/${TAG}_bench_mufast.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_mufast.json | cut -c1-250
