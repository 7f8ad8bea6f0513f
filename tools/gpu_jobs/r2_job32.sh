#!/bin/bash
# round-2 GPU job 32: next epoch's tuples packed beside this epoch's SGD kernel (lfm_plan_epoch_next)
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_api.py tests/test_gpu_tierb.py tests/test_gpu_hogwild.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_pytest32.log
tail -2 gpurun_out/r2_pytest32.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-c4 --no-ranks --no-replay --no-traffic --no-cpu-baseline > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
tail -2 gpurun_out/r2_bench32.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench32.json')); e=b['e2e']; print(b['value'], b['ms_per_step'], b['detail']['wall_ms_per_step_resident'], b['roofline']['kernel_ms'], e['value'], e['ms_per_step'], e['five_epochs_interactions_per_s'], b['detail']['weights_finite'])"
echo job32 done
