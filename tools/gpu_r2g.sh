#!/bin/bash
TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_decomposed.py -m gpu -q > $O/${TAG}_pytest_decomposed.log 2>&1; echo "decomposed pytest exit $?"; tail -8 $O/${TAG}_pytest_decomposed.log
timeout 600 python tools/dropin_bench.py --n 128 256 --steps 20 --out $O/${TAG}_dropin.json > $O/${TAG}_dropin.log 2>&1; tail -3 $O/${TAG}_dropin.log
