#!/bin/bash
TAG=${1:-r02h}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $OUT/pytest_gpu_$TAG.log; tail -4 $OUT/pytest_gpu_$TAG.log
for v in "" "--egroups 64" "--geometry" "--egroups 64 --regions-2d 10" "--regions-2d 320000" "--exp mufu"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs $v 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('[$v] %.4e int/s  %.3f ms  %s frac %.3f  e2e %.4e  sm %s  %s'%(d['value'],d['ms_per_step'],d['roofline']['bound'],d['roofline']['frac'],d['e2e']['value'],d['clocks']['sm_mhz'],d['roofline']['kernel'][-40:]))" | tee -a $OUT/bench_quick_$TAG.txt
done
SMK_ADDR64=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs 2>&1 | tail -1 | cut -c1-200
