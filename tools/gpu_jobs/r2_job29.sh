#!/bin/bash
# round-2 GPU job 29: final evidence -- full suite, smoke, bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest29.log
tail -3 gpurun_out/r2_pytest29.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke29.log 2>&1; tail -2 gpurun_out/r2_smoke29.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench29.json 2> gpurun_out/r2_bench29.err
tail -2 gpurun_out/r2_bench29.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench29.json')); e=b['e2e']; print(b['value'], e['value'], e['ms_per_step'], e['cold']['value'], e['cold']['ms_per_step'], b['roofline']['kernel_ms'], b['roofline']['traffic']); print(b['parity']['p_at_10_gpu'], b['parity']['p_at_10_reference'], b['parity']['heldout_auc_gpu'], b['parity']['heldout_auc_reference'], b['cpu_baseline']['value']); print({k:(v['interactions_per_s_kernel'], v['cpu_baseline']['value']) for k,v in b['replay'].items()}); print(b['c4']['single']['interactions_per_s']); print({k:(v.get('kernel_ms'),v.get('call_wall_ms')) for k,v in b['ranks'].items() if isinstance(v,dict)})"
echo job29 done
