#!/bin/bash
# round-2 GPU job 6: reference-suite order diagnostic, full suite, C3 v5, ncu captures of the final kernels
# (ncu reports are summarised ON the box and deleted: gpurun copies back at most 64 MiB)
mkdir -p gpurun_out
summarise() {  # summarise <rep-basename> <interactions per launch> : raw metrics csv + summary, then drop the report
  python tools/ncu_summary.py gpurun_out/$1.ncu-rep $2 20 > gpurun_out/$1_summary.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 400 python tools/diag_refsuite.py > gpurun_out/r2_diag_refsuite.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_pytest6.log
timeout 300 python tools/bench_c3.py 20000000 hot 2 > gpurun_out/r2_c3e.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fast_slot_kernel -s 3 -c 1 \
    -o gpurun_out/r2_c2_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c4 --no-ranks \
    > gpurun_out/r2_c2_final_ncu.log 2>&1
summarise r2_c2_final 20e6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4 --no-ranks > gpurun_out/r2_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_ranks_tiled -c 1 \
    -o gpurun_out/r2_ranks_g3 -f python tools/bench_ranks.py 20000 3 > gpurun_out/r2_ranks_ncu3.log 2>&1
summarise r2_ranks_g3 2e9
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hogwild_kernel -c 1 --launch-skip 1 \
    -o gpurun_out/r2_c3_hot_v5 -f python tools/bench_c3.py 4000000 hot 1 > gpurun_out/r2_c3_ncu_v5.log 2>&1
summarise r2_c3_hot_v5 4e6
du -sh gpurun_out
echo job6 done
