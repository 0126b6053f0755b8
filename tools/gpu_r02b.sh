#!/bin/bash
# round 2, second GPU call: tests, smoke, bench, expf sweep, ncu launch list + --set full captures of every hot kernel
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_gpu_$TAG.log; tail -5 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cut -c1-400 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== expf sweep"; timeout 900 python tools/expf_sweep.py > $OUT/expf_sweep_$TAG.md 2>&1; tail -4 $OUT/expf_sweep_$TAG.md
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-legs > $OUT/ncu_launch_bench_$TAG.log 2>&1
cap() {  # name, bench args
  local name=$1; shift
  echo "== ncu full $name"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_${name}_$TAG \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-legs "$@" > $OUT/ncu_full_${name}_$TAG.log 2>&1
  tail -1 $OUT/ncu_full_${name}_$TAG.log | cut -c1-200
}
cap default
cap hbm --regions-2d 320000
cap g7 --egroups 7
cap g64c4 --egroups 64 --regions-2d 10
cap geom --geometry
ls -la $OUT | tail -30
