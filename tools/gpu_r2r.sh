#!/bin/bash
# round 2, final 1-GPU visit: the whole GPU suite, smoke(), the bench line exactly as the driver runs it, the reference arm,
# the ncu launch list of the bench command
TAG=${1:-r02r}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -n 4 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
/usr/bin/time -v timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-250; grep "Elapsed (wall" $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench_ref.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-sub > $O/${TAG}_ncu_bench.log 2>&1
