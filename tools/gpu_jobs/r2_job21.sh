#!/bin/bash
# round-2 GPU job 21: async model upload + pack on a side stream (public-API e2e); five-config coverage table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_tierb.py tests/test_gpu_hogwild.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest21.log
tail -3 gpurun_out/r2_pytest21.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-c4 --no-ranks --no-replay --no-traffic --no-cpu-baseline > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench21.json')); e=b['e2e']; print(b['value'], e['value'], e['ms_per_step'], e['ms_per_step_min_max'], e['five_epochs_interactions_per_s'], e['cold']['value'])"
timeout 600 python tools/bench_configs.py C1,C3,C4s,C5 > gpurun_out/r2_configs21.jsonl 2> gpurun_out/r2_configs21.err
cat gpurun_out/r2_configs21.jsonl | cut -c1-330
echo job21 done
