#!/bin/bash
# round-2 GPU job 12: fused dataflow replay kernel (scheduler warp + executors)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py tests/test_gpu_replay_parity.py tests/test_gpu_replay_edge_cases.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest12_dataflow.log
tail -5 gpurun_out/r2_pytest12_dataflow.log
timeout 400 python tools/bench_replay.py C1,logistic,C2-shape-bpr,C5-slice-logistic > gpurun_out/r2_replay12.jsonl 2> gpurun_out/r2_replay12.err
cat gpurun_out/r2_replay12.jsonl; tail -3 gpurun_out/r2_replay12.err
echo job12 done
