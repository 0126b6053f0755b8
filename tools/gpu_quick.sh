#!/bin/bash
# quick GPU check: full GPU test-suite + smoke + one bench line
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$TAG.json
