#!/bin/bash
# round-2 GPU job 30: sampler resolves the lanes event by event
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py tests/test_gpu_replay_parity.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest30.log
tail -2 gpurun_out/r2_pytest30.log
LFM_RDF_PROFILE=1 timeout 200 python tools/bench_replay.py C1,C2-shape-bpr > gpurun_out/r2_replay30.jsonl 2> gpurun_out/r2_replay30.err
cut -c1-160 gpurun_out/r2_replay30.jsonl; grep rdf gpurun_out/r2_replay30.err
echo job30 done
