#!/bin/bash
# round-2 GPU job 2: full GPU suite again (after the top-k fix), replay / C3 / C2-variant timings
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -200 > gpurun_out/r2_pytest2.log
timeout 300 python tools/bench_replay.py > gpurun_out/r2_replay.log 2>&1
timeout 600 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3b.log 2>&1
timeout 600 python tools/tune_warp.py 7,9,10,6 > gpurun_out/r2_tune.log 2>&1
echo job2 done
