#!/bin/bash
# round-2 GPU job 15: final dataflow scheduler; whole GPU suite; bench with the replay block
mkdir -p gpurun_out
LFM_RDF_PROFILE=1 timeout 300 python tools/bench_replay.py C1,logistic,C2-shape-bpr,C5-slice-logistic > gpurun_out/r2_replay15.jsonl 2> gpurun_out/r2_replay15.err
cat gpurun_out/r2_replay15.jsonl; grep rdf gpurun_out/r2_replay15.err
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest15.log
tail -3 gpurun_out/r2_pytest15.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err
tail -3 gpurun_out/r2_bench15.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench15.json')); print(b['value'], b['e2e']['value']); print(json.dumps(b['replay'])[:3000])"
echo job15 done
