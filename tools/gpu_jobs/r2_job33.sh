#!/bin/bash
# round-2 GPU job 33: the bench line at HEAD (last GPU minutes of the round)
mkdir -p gpurun_out
timeout 170 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench33.json 2> gpurun_out/r2_bench33.err
tail -2 gpurun_out/r2_bench33.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench33.json')); e=b['e2e']; print(b['value'], e['value'], e['cold']['value'], b['roofline']['traffic']); print({k:(v['interactions_per_s_kernel'], v['cpu_baseline']['value']) for k,v in b['replay'].items()}); print(b['c4']['single']['interactions_per_s'] if b['c4'] and 'single' in b['c4'] else b['c4'])"
echo job33 done
