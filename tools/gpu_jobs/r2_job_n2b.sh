#!/bin/bash
# round-2 GPU job on TWO GPUs (re-entry): 2-rank sharded-fit quality tests, bench.py --gpus 2 (weak-scaled C2 + C4 both axes)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n2_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2_pytest_n2.log
tail -4 gpurun_out/r2_pytest_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
cut -c1-300 gpurun_out/r2_bench_n2.json
echo job_n2 done
