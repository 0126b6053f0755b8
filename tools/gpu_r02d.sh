#!/bin/bash
TAG=${1:-r02d}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $OUT/pytest_gpu_$TAG.log; tail -4 $OUT/pytest_gpu_$TAG.log
bash tools/gpu_ab.sh ${TAG}_g7 "--egroups 7" main noglibc
bash tools/gpu_ab.sh ${TAG}_g29 "--egroups 29" main noglibc
bash tools/gpu_ab.sh ${TAG}_g128 "" main noglibc
