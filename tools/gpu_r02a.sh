#!/bin/bash
# round 2, first GPU call: the whole GPU test suite, smoke, the bench line, a first variant comparison
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
(lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; nproc) > $OUT/host_$TAG.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 900 python bench.py 2>$OUT/bench_$TAG.err | tail -1 | tee $OUT/bench_$TAG.json | cut -c1-600
tail -5 $OUT/bench_$TAG.err
for lib in "" simplemoc-kernel_b200/lib/variants/libsmk_mb5.so; do
  for i in 1 2; do
    echo "== variant '$lib' run $i"
    SMK_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4e  e2e %.4e  ms %.3f  frac %.3f  clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" | tee -a $OUT/variants_$TAG.txt
  done
done
