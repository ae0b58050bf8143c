#!/bin/bash
# Round-2 (hybrid sparse + dense-band pile-up) profiling pass (GPU box, under gpurun): launch list of a bench step and
# full ncu captures of the two pile-up kernels on the chr1 launch of configs[3].
set -x
M=gpu__time_duration.sum
ncu --metrics $M --clock-control none -s 700 -c 1400 --csv --log-file gpurun_out/r2c_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2c_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pileup_main -s 24 -c 1 -o gpurun_out/r2c_prof_main_c3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2c_prof_main.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pileup_dense -s 24 -c 1 -o gpurun_out/r2c_prof_dense_c3 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2c_prof_dense.log 2>&1
for w in main dense; do
  ncu -i gpurun_out/r2c_prof_${w}_c3.ncu-rep --page raw --csv > gpurun_out/r2c_prof_${w}_c3.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
