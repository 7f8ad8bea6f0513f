#!/bin/bash
# round-2 GPU job 20: evidence for HEAD -- full suite, bench line, ncu of the (changed) C2 kernel, launch list
mkdir -p gpurun_out
summarise() {
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep $2 20 > gpurun_out/$1_summary.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep
}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_pytest20.log
tail -3 gpurun_out/r2_pytest20.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
tail -2 gpurun_out/r2_bench20.err
python -c "
import json; b=json.load(open('gpurun_out/r2_bench20.json')); print(b['value'], b['e2e']['value'], b['roofline']['kernel_ms'], b['roofline']['traffic']); print({k:(v['interactions_per_s_kernel'], v['cpu_baseline']['value']) for k,v in b['replay'].items()}); print({k:(v.get('kernel_ms'),v.get('call_wall_ms')) for k,v in b['ranks'].items() if isinstance(v,dict)})"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_slot_kernel -s 3 -c 1 \
    -o gpurun_out/r2_c2_final2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-ranks --no-replay --no-traffic \
    > gpurun_out/r2_c2_final2_ncu.log 2>&1
summarise r2_c2_final2 20e6
head -24 gpurun_out/r2_c2_final2_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-ranks --no-traffic > gpurun_out/r2_launches_bench_final.log 2>&1
grep -c rdf_kernel gpurun_out/r2_launches_bench_steps2_final.csv
echo job20 done
