#!/bin/bash
# round-2 GPU job 28: predict_ranks with two user tiles per CTA (128 registers, unroll 8)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_eval_topk.py tests/test_gpu_scoring.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest28.log
tail -2 gpurun_out/r2_pytest28.log
for g in 3 2 1; do timeout 200 python tools/bench_ranks.py 20000 $g >> gpurun_out/r2_ranks28.jsonl 2>> gpurun_out/r2_ranks28.err; done
cut -c1-200 gpurun_out/r2_ranks28.jsonl
echo job28 done
