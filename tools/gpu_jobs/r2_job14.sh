#!/bin/bash
# round-2 GPU job 14: scheduler with runtime window + mark arrays
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py tests/test_gpu_replay_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest14_dataflow.log
tail -3 gpurun_out/r2_pytest14_dataflow.log
LFM_RDF_PROFILE=1 timeout 300 python tools/bench_replay.py C1,logistic,C2-shape-bpr,C5-slice-logistic > gpurun_out/r2_replay14.jsonl 2> gpurun_out/r2_replay14.err
cat gpurun_out/r2_replay14.jsonl; grep rdf gpurun_out/r2_replay14.err
echo job14 done
