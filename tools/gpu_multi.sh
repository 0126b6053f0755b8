#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus N): the multi-GPU tests, then bench.py under torchrun at N ranks
N=${1:-2}; TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi_multi${N}_$TAG.csv 2>&1
echo "== pytest -m gpu -k multi"; timeout 900 python -m pytest tests -m gpu -q -k "multi or driver" 2>&1 | tail -15 | tee $OUT/pytest_multi${N}_$TAG.log
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err
tail -c 2500 $OUT/bench_n${N}_$TAG.json; tail -5 $OUT/bench_n${N}_$TAG.err
echo "== reference arm N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 | cut -c1-600 | tee $OUT/bench_ref_n${N}_$TAG.json
echo "== C driver --gpus $N"
timeout 300 ./simplemoc-kernel_b200/bin/SimpleMOC-kernel -s 1000000000 --gpus $N 2>&1 | tail -25 | tee $OUT/driver_n${N}_$TAG.log
