#!/bin/bash
# round 2: full GPU suite after the comm / Roe / characteristic changes + default bench (with all sub-records)
TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_roe_fused.py -m gpu -q > $O/${TAG}_pytest_functors.log 2>&1; echo "functors pytest exit $?"; tail -4 $O/${TAG}_pytest_functors.log
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -1 $O/${TAG}_bench.json | cut -c1-300; tail -3 $O/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
