#!/bin/bash
# perf exploration: tests + bench for alternative library builds + ncu of the default build
TAG=${1:-p}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/pytest_gpu_$TAG.log
for lib in libsmk.so libsmk_mb3.so; do
  for v in "" "--exp mufu" "--egroups 64" "--egroups 7" "--egroups 64 --regions-2d 10"; do
    echo "== bench $lib $v"; SMK_LIB=$PWD/simplemoc-kernel_b200/lib/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('%.4e int/s  %.3f ms  frac %.3f  e2e %.4e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e: print('ERR',l[:300])
" | tee -a $OUT/bench_variants_$TAG.txt
  done
done
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline > $OUT/ncu_full_bench_$TAG.log 2>&1
ls -la $OUT | tail -5
