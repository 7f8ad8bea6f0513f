#!/bin/bash
# round-2 GPU job 13: where does the scheduler warp's time go (phase cycle counters), and how much the executors' polling costs it
mkdir -p gpurun_out
export LFM_RDF_PROFILE=1
for ctas in 148 40 12; do
  echo "== LFM_RDF_CTAS=$ctas" >> gpurun_out/r2_replay13.err
  LFM_RDF_CTAS=$ctas timeout 300 python tools/bench_replay.py C1,C2-shape-bpr,C5-slice-logistic >> gpurun_out/r2_replay13.jsonl 2>> gpurun_out/r2_replay13.err
done
cat gpurun_out/r2_replay13.err | grep -v "^$" | tail -40
echo job13 done
