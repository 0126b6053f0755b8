#!/bin/bash
# gather-record kernel (<= 32 groups): parity tests + A/B against the general kernel on ONE box.
#   SMK_RECORDS=0 general kernel | 2, 4 = record kernel with that many groups per lane
#   extra libs (lib/variants/libsmk_<name>.so) can be listed as arguments: run with SMK_RECORDS=2
TAG=${TAG:-rec}
OUT=gpurun_out; mkdir -p $OUT
echo "== record tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "record or fast_mode_within or strict_mode or row_range or zero_and" 2>&1 | tail -8 | tee $OUT/pytest_$TAG.log
run() {  # label, lib, SMK_RECORDS, args
  SMK_LIB=$2 SMK_RECORDS=$3 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-legs $4 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1 [$4] : %.4e int/s  %.3f ms  frac %.3f  e2e %.4e  sm %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['clocks']['sm_mhz']))
except Exception as e: print('$1 ERR',l[:300])
" | tee -a $OUT/ab_$TAG.txt
}
MAIN=$PWD/simplemoc-kernel_b200/lib/libsmk.so
for rep in 1 2; do
for G in 7 29 13 3; do
  for m in 0 2 4; do run "records=$m" $MAIN $m "--egroups $G"; done
  for v in "$@"; do run "$v(records=2)" $PWD/simplemoc-kernel_b200/lib/variants/libsmk_$v.so 2 "--egroups $G"; done
done; done
echo "== ncu g7 record kernel"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attenuate_record -c 1 -o $OUT/prof_g7rec_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-legs --egroups 7 > $OUT/ncu_g7rec_$TAG.log 2>&1
python tools/ncu_summary.py $OUT/prof_g7rec_$TAG.ncu-rep --title "g7 record kernel ($TAG)" --intersections 7e8 > $OUT/ncu_g7rec_$TAG.md 2>&1 || tail -5 $OUT/ncu_g7rec_$TAG.md
tail -30 $OUT/ncu_g7rec_$TAG.md
