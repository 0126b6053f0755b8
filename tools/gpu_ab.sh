#!/bin/bash
# A/B of library builds on ONE box: tools/gpu_ab.sh <tag> "<bench args>" lib1 lib2 ...   (lib = name under lib/variants, or "main")
TAG=$1; ARGS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = main ]; then L=$PWD/simplemoc-kernel_b200/lib/libsmk.so; else L=$PWD/simplemoc-kernel_b200/lib/variants/libsmk_$lib.so; fi
  SMK_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs $ARGS 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$lib [$ARGS] rep$rep: %.4e int/s  %.3f ms  frac %.3f  e2e %.4e  sm %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['clocks']['sm_mhz']))
except Exception as e: print('$lib ERR',l[:300])
" | tee -a $OUT/ab_$TAG.txt
done; done
