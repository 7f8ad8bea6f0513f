#!/bin/bash
# round-2 GPU job 7: rank diagnostic (300 random models vs oracle), suite, C3 v6, bench
mkdir -p gpurun_out
timeout 600 python tools/diag_refsuite.py > gpurun_out/r2_diag_refsuite2.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_pytest7.log
timeout 300 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3f.log 2>&1
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
du -sh gpurun_out
echo job7 done
