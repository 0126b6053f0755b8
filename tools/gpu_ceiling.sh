#!/bin/bash
# timing experiments: the default kernel with one ingredient changed (values are wrong in these builds)
OUT=gpurun_out
for lib in libsmk.so libsmk_NO_MUFU.so libsmk_NO_RED.so libsmk_ONE_TYPE.so; do
  echo -n "$lib: "; SMK_LIB=$PWD/simplemoc-kernel_b200/lib/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('%.4e int/s  %.3f ms'%(d['value'],d['ms_per_step']))
except Exception as e: print('ERR',l[:300])
"
  SMK_LIB=$PWD/simplemoc-kernel_b200/lib/$lib timeout 300 ncu --metrics smsp__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:attenuate -s 1 -c 1 python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline 2>&1 | grep -E "inst_executed|fma_cycles|issue_active|duration" | awk '{print "    ",$1,$2,$3}'
done
