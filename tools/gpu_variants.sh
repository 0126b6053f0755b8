#!/bin/bash
# compare kernel variants (SMK_KERNEL knob): parity tests under each + bench
TAG=${1:-v}; shift
VARIANTS=${@:-direct prefetch staged2 staged3}
OUT=gpurun_out
mkdir -p $OUT
for k in $VARIANTS; do
  echo "== SMK_KERNEL=$k pytest"; SMK_KERNEL=$k timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
  for v in "" "--exp mufu"; do
    echo "== bench $k $v"; SMK_KERNEL=$k timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('%.4e int/s  %.3f ms  frac %.3f  e2e %.4e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']))
except Exception as e: print('ERR',l[:300])
" | tee -a $OUT/bench_variants_$TAG.txt
  done
done
K=$(echo $VARIANTS | awk '{print $NF}')
echo "== ncu full $K"
SMK_KERNEL=$K timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline > $OUT/ncu_full_bench_$TAG.log 2>&1
tail -2 $OUT/ncu_full_bench_$TAG.log | cut -c1-200
