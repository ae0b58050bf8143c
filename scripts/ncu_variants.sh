#!/bin/bash
# ncu counters of the main kernel on the chr1 launch of the bench workload, one run per tuning variant (GPU box).
# usage: scripts/ncu_variants.sh OUT_PREFIX "ENV1=.. ENV2=.." "ENV..." ...
out=$1; shift
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size
i=0
for v in "$@"; do
  echo "== variant $i: $v"
  env $v ncu --metrics $M --clock-control none -k regex:k_pileup_main -s 1 -c 1 --csv --log-file ${out}_$i.csv \
      python bench.py --chroms chr1 --steps 1 --warmup 1 --no-cpu --no-e2e > ${out}_$i.log 2>&1
  echo "variant: $v" >> ${out}_$i.csv
  i=$((i+1))
done
