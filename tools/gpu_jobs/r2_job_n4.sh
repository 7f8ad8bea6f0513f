#!/bin/bash
# round-2 GPU job on FOUR GPUs: bench.py --gpus 4 (weak-scaled C2 + C4 strong scaling on both axes)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_n4_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29548 \
    bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -3 gpurun_out/r2_bench_n4.err
grep "^{" gpurun_out/r2_bench_n4.json | cut -c1-200
echo job_n4 done
