#!/bin/bash
# round-2 GPU job 1: full GPU test suite, C3 feature path (hot rows on/off) + ncu, headline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_pytest1.log
timeout 600 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3.log 2>&1
for v in hot nohot; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:hogwild_kernel -c 1 --launch-skip 1 \
    -o gpurun_out/r2_c3_$v -f python tools/bench_c3.py 4000000 $v 1 > gpurun_out/r2_c3_ncu_$v.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo job1 done
