#!/bin/bash
# round 2, first GPU call: tests, smoke, default bench line, reference arm
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
(lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; nproc) > $OUT/host_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference_$TAG.json
