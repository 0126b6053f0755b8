#!/bin/bash
# ceiling experiments: the same kernel without tallies / without loads / without both
for lib in libsmk.so libsmk_NO_RED.so libsmk_NO_LDG.so libsmk_NO_BOTH.so; do
  for k in direct flat; do
    echo -n "$lib $k: "; SMK_KERNEL=$k SMK_LIB=$PWD/simplemoc-kernel_b200/lib/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('%.4e int/s  %.3f ms'%(d['value'],d['ms_per_step']))
except Exception as e: print('ERR',l[:300])
"
  done
done
