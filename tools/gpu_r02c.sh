#!/bin/bash
# round 2, third GPU call: tests on the new arithmetic, A/B of timing variants, ncu captures summarised ON THE BOX
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/pytest_gpu_$TAG.log; tail -5 $OUT/pytest_gpu_$TAG.log
echo "== A/B"
bash tools/gpu_ab.sh $TAG "" main pfsigt noglibc pfsigt_noglibc
bash tools/gpu_ab.sh ${TAG}_g64 "--egroups 64" main pfsigt
echo "== bench (default, all legs)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cut -c1-300 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs > $OUT/ncu_launch_bench_$TAG.log 2>&1
cap() {  # name, intersections, bench args
  local name=$1; local inter=$2; shift 2
  echo "== ncu full $name"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_${name}_$TAG \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-legs "$@" > $OUT/ncu_full_${name}_$TAG.log 2>&1
  python tools/ncu_summary.py $OUT/prof_${name}_$TAG.ncu-rep --title "round 2 ($TAG): $name -- bench.py $*" --intersections $inter > $OUT/ncu_${TAG}_$name.md 2> $OUT/ncu_summary_${name}.err
  ncu -i $OUT/prof_${name}_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > $OUT/src_${name}_$TAG.csv.gz
  [ "$name" = default ] || rm -f $OUT/prof_${name}_$TAG.ncu-rep
  head -12 $OUT/ncu_${TAG}_$name.md | tail -6
}
cap default 12800000000
cap hbm 12800000000 --regions-2d 320000
cap g7 700000000 --egroups 7
cap g64c4 6400000000 --egroups 64 --regions-2d 10
cap geom 12800000000 --geometry
du -sh $OUT; ls -la $OUT | tail -30
