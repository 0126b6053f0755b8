#!/bin/bash
OUT=gpurun_out
for v in NO_RED NO_LDG; do
  SMK_LIB=$PWD/simplemoc-kernel_b200/lib/libsmk_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_$v \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline > $OUT/ncu_$v.log 2>&1
  tail -1 $OUT/ncu_$v.log | cut -c1-100
done
