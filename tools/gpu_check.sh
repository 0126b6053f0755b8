#!/bin/bash
# One gpurun call: GPU tests, expf sweep, bench (+ variants), ncu launch list and full capture.
# usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.csv 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/host_$TAG.txt; nproc >> $OUT/host_$TAG.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_$TAG.json
for v in "--exp mufu" "--exp glibc" "--math strict --exp glibc" "--exp table" "--egroups 7" "--egroups 64 --regions-2d 10" "--egroups 64"; do
  echo "== bench $v"; timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $v 2>&1 | tail -1 | tee -a $OUT/bench_variants_$TAG.jsonl
done
echo "== driver"; ./simplemoc-kernel_b200/bin/SimpleMOC-kernel 2>&1 | tail -22 | tee $OUT/driver_$TAG.log
echo "== expf sweep"; timeout 600 python tools/expf_sweep.py > $OUT/expf_sweep_$TAG.md 2>&1; tail -4 $OUT/expf_sweep_$TAG.md
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_bench_$TAG.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attenuate -s 1 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 1 --warmup 1 --segments 20000000 --no-cpu-baseline > $OUT/ncu_full_bench_$TAG.log 2>&1
ls -la $OUT
