#!/bin/bash
# round-2 GPU job 17: membership words fetched with the candidate rows (slot kernels); parity tests of the hogwild tier; bench with in-run traffic
mkdir -p gpurun_out
timeout 300 python tools/tune_warp.py 9,7 > gpurun_out/r2_tune17.jsonl 2> gpurun_out/r2_tune17.err
cat gpurun_out/r2_tune17.jsonl
timeout 900 python -m pytest tests/test_gpu_hogwild.py tests/test_gpu_probe.py tests/test_gpu_tierb.py tests/test_gpu_fullsize.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest17.log
tail -3 gpurun_out/r2_pytest17.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err
tail -3 gpurun_out/r2_bench17.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench17.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline']['kernel_ms'], b['roofline']['traffic'], b['roofline']['traffic_source'][:60]); print(b['parity']); print(b['c4']['single']['interactions_per_s']); print({k:v.get('call_wall_ms') for k,v in b['ranks'].items() if isinstance(v,dict)})"
echo job17 done
