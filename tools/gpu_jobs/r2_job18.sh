#!/bin/bash
# round-2 GPU job 18: three-stage scheduler pipeline of the dataflow replay; top-k select with warp-aggregated histogram
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_replay_dataflow.py tests/test_gpu_replay_parity.py tests/test_gpu_eval_topk.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_pytest18.log
tail -3 gpurun_out/r2_pytest18.log
LFM_RDF_PROFILE=1 timeout 300 python tools/bench_replay.py C1,logistic,C2-shape-bpr,C5-slice-logistic > gpurun_out/r2_replay18.jsonl 2> gpurun_out/r2_replay18.err
cat gpurun_out/r2_replay18.jsonl; grep rdf gpurun_out/r2_replay18.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-c4 --no-replay --no-traffic --no-cpu-baseline > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench18.json')); print({k:(v.get('kernel_ms'),v.get('call_wall_ms')) for k,v in b['ranks'].items() if isinstance(v,dict)})"
echo job18 done
