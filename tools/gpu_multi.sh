#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus N): [the multi-GPU tests,] bench.py under torchrun at N ranks, the C driver
N=${1:-2}; TAG=${2:-r02}; TESTS=${3:-yes}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi_multi${N}_$TAG.csv 2>&1
if [ "$TESTS" = yes ]; then
  echo "== pytest -m gpu -k multi"; timeout 900 python -m pytest tests -m gpu -q -k "multi or driver" 2>&1 | tail -15 | tee $OUT/pytest_multi${N}_$TAG.log
fi
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
tail -c 1500 $OUT/bench_n${N}_$TAG.json; tail -5 $OUT/bench_n${N}_$TAG.err | grep -v "^\*\|OMP_NUM\|^$"
echo "== C driver --gpus $N (config 5: 1e10 segments)"
timeout 300 ./simplemoc-kernel_b200/bin/SimpleMOC-kernel -s 10000000000 --gpus $N 2>&1 | tail -16 | tee $OUT/driver_n${N}_$TAG.log
