#!/bin/bash
# End-of-round GPU evidence in one gpurun call: tests, smoke, bench line, per-config bench lines,
# ncu launch list and ncu --set full capture of the default kernel.  Outputs under gpurun_out/.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
(lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; nproc) > $OUT/host_$TAG.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench (default)"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
echo "== bench (reference arm)"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference_$TAG.json
: > $OUT/bench_configs_$TAG.jsonl
for v in "--exp mufu" "--exp glibc" "--exp table" "--math strict --exp glibc" \
         "--egroups 7" "--egroups 64 --regions-2d 10" "--egroups 64" "--egroups 256 --segments 50000000" \
         "--regions-2d 320000" "--segments 10000000000 --steps 1 --warmup 1"; do
  echo "== bench $v"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $v 2>&1 | tail -1 | tee -a $OUT/bench_configs_$TAG.jsonl | cut -c1-160
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench_$TAG.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline > $OUT/ncu_full_bench_$TAG.log 2>&1
echo "== ncu full, HBM-resident regime"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_hbm_$TAG \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --regions-2d 320000 --no-cpu-baseline > $OUT/ncu_full_hbm_$TAG.log 2>&1
ls -la $OUT | tail -20
