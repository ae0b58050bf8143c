#!/bin/bash
# launch list of our kernels only (torch's genome generation filtered out), per-chromosome algorithmic bytes,
# compute-sanitizer memcheck of the kernels touched in round 2
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_|cub::' -s 330 -c 1400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-api > gpurun_out/r2_launches_bench.log 2>&1
PUP_BENCH_VERBOSE=1 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-api 2> gpurun_out/r2_cost_c3.json > /dev/null
PUP_BENCH_VERBOSE=1 python bench.py --workload configs4 --steps 1 --warmup 1 --no-cpu --no-e2e --no-api 2> gpurun_out/r2_cost_c4.json > /dev/null
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_windows.py -m gpu -x -q \
    -k "variants and 300-21 or adversarial and 2-8 or control_shifts_many or equal_host_windows and toy_strand_dist_ctrl or golden_case_through_cuda and trans_ctrl or async_upload" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo memcheck rc=$?
tail -5 gpurun_out/r2_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_device_windows.py -m gpu -x -q -k "control_shifts_other or equal_host_windows and toy_bywindow" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo racecheck rc=$?
tail -5 gpurun_out/r2_sanitizer_racecheck.log
