#!/bin/bash
# round-2 GPU job 26: single-pass top-k for k <= 16
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_eval_topk.py tests/test_gpu_reference_suite.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest26.log
tail -3 gpurun_out/r2_pytest26.log
timeout 300 python tools/bench_recommend.py > gpurun_out/r2_recommend26.json 2> gpurun_out/r2_recommend26.err
cat gpurun_out/r2_recommend26.json
timeout 300 python tools/bench_recommend.py 20000 100 >> gpurun_out/r2_recommend26.json 2>> gpurun_out/r2_recommend26.err
tail -1 gpurun_out/r2_recommend26.json
echo job26 done
