#!/bin/bash
# End-of-round GPU evidence in ONE gpurun call: tests, smoke, bench (all legs), reference arm, expf sweep, compute-sanitizer,
# ncu launch list and ncu --set full captures of every hot kernel, summarised ON THE BOX (tools/ncu_summary.py) so that only
# the markdown summaries, the per-instruction source pages (gz) and one .ncu-rep travel back (gpurun_out/ is capped at 64 MiB).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
(lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; nproc) > $OUT/host_$TAG.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/pytest_gpu_$TAG.log; tail -4 $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cut -c1-300 $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_reference_$TAG.json; cut -c1-200 $OUT/bench_reference_$TAG.json
: > $OUT/bench_modes_$TAG.jsonl
for v in "--exp mufu" "--exp glibc" "--exp table" "--math strict --exp glibc" "--egroups 64" "--egroups 256 --segments 50000000" "--egroups 29"; do
  echo "== bench $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-legs $v 2>&1 | tail -1 | tee -a $OUT/bench_modes_$TAG.jsonl | cut -c1-160
done
echo "== expf sweep"; timeout 900 python tools/expf_sweep.py > $OUT/expf_sweep_$TAG.md 2>&1; tail -6 $OUT/expf_sweep_$TAG.md
echo "== compute-sanitizer"
( for tool in memcheck racecheck synccheck; do
    echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or record or fit_per_sweep or pipelined or (geometry_strict and 200-5-128)" 2>&1 | tail -4
  done ) > $OUT/sanitizer_$TAG.log 2>&1; grep -E "SUMMARY|passed|failed" $OUT/sanitizer_$TAG.log
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
  sed -n 7,8p $OUT/ncu_${TAG}_$name.md
}
cap default 12800000000
cap hbm 12800000000 --regions-2d 320000
cap g7 700000000 --egroups 7
cap g64c4 6400000000 --egroups 64 --regions-2d 10
cap geom 12800000000 --geometry
du -sh $OUT
