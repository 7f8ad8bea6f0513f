#!/bin/bash
# round-2 GPU job 23: version stage requests the next chunk's counters early
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py tests/test_gpu_replay_parity.py tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest23.log
tail -3 gpurun_out/r2_pytest23.log
LFM_RDF_PROFILE=1 timeout 300 python tools/bench_replay.py C1,logistic,C2-shape-bpr,C5-slice-logistic > gpurun_out/r2_replay23.jsonl 2> gpurun_out/r2_replay23.err
cut -c1-200 gpurun_out/r2_replay23.jsonl; grep rdf gpurun_out/r2_replay23.err
echo job23 done
