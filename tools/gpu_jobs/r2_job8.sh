#!/bin/bash
# round-2 GPU job 8 (re-entry): state of HEAD -- full GPU suite, C3 hot/nohot, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -200 > gpurun_out/r2_pytest8.log
tail -5 gpurun_out/r2_pytest8.log
timeout 300 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3_job8.log 2>&1
tail -3 gpurun_out/r2_c3_job8.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -3 gpurun_out/r2_bench8.err
cut -c1-400 gpurun_out/r2_bench8.json
echo job8 done
