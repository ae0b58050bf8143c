#!/bin/bash
# Round-2 profiling pass (GPU box, under gpurun): launch list of a bench step, full ncu captures of the dominant
# kernel for configs[3] (W = 83) and configs[4] (W = 203), summaries as CSV for profiles/.
set -x
M=gpu__time_duration.sum
ncu --metrics $M --clock-control none -s 700 -c 800 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pileup_main -s 24 -c 1 -o gpurun_out/r2_prof_main_c3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2_prof_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pileup_main -s 24 -c 1 -o gpurun_out/r2_prof_main_c4 -f \
    python bench.py --workload configs4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2_prof_c4.log 2>&1
for w in c3 c4; do
  ncu -i gpurun_out/r2_prof_main_$w.ncu-rep --page raw --csv > gpurun_out/r2_prof_main_$w.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
