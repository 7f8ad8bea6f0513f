#!/bin/bash
# round-2 GPU job 19: scheduler pipeline -- what the stage hand-off costs (fences / spinning)
mkdir -p gpurun_out
for f in 0 1 2 3; do
  echo "== LFM_RDF_FLAGS=$f" >> gpurun_out/r2_replay19.err
  LFM_RDF_FLAGS=$f LFM_RDF_PROFILE=1 timeout 300 python tools/bench_replay.py C1,C5-slice-logistic >> gpurun_out/r2_replay19.jsonl 2>> gpurun_out/r2_replay19.err
done
grep "rdf\|==" gpurun_out/r2_replay19.err
echo job19 done
