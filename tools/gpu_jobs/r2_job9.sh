#!/bin/bash
# round-2 GPU job 9: dataflow replay (new) first, then the state of the whole suite, C3, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_pytest9_dataflow.log
tail -5 gpurun_out/r2_pytest9_dataflow.log
timeout 400 python tools/bench_replay.py > gpurun_out/r2_replay9.jsonl 2> gpurun_out/r2_replay9.err
cat gpurun_out/r2_replay9.jsonl
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_replay_dataflow.py 2>&1 | tail -200 > gpurun_out/r2_pytest9.log
tail -5 gpurun_out/r2_pytest9.log
timeout 300 python tools/bench_c3.py 20000000 hot,nohot 2 > gpurun_out/r2_c3_job9.log 2>&1
tail -3 gpurun_out/r2_c3_job9.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
tail -3 gpurun_out/r2_bench9.err
cut -c1-400 gpurun_out/r2_bench9.json
echo job9 done
