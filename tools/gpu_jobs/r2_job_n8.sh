#!/bin/bash
# round-2 GPU job on EIGHT GPUs: bench.py --gpus 8 (weak-scaled C2 + C4 strong scaling on both axes)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_n8_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
grep "^{" gpurun_out/r2_bench_n8.json | cut -c1-200
echo job_n8 done
