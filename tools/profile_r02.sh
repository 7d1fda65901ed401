# round-2 ncu evidence (run under gpurun, one GPU); summaries are made on the CPU box from the files this leaves in gpurun_out/
set -x
M=gpu__time_duration.sum
# (1) launch list of the bench command's timed region (one step of the 1024^3 workload, every launch, cold-cache and serialised)
CFB_NCU_RANGE=1 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_bench_1024.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-pageable --no-alt-precision > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
# (2) --set full of every network kernel, 4 patches per launch
CFB_PROFILE_CHUNK=32,448,448 CFB_BENCH_BATCH=4 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'conv3|convT|first_conv|maxpool|normalize' -o gpurun_out/r02_prof_full python tools/profile_step.py > gpurun_out/r02_prof_full.log 2>&1
# (3) the HBM-bound kernels
ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r02_memory_kernels.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  python tools/profile_memory_kernels.py > gpurun_out/r02_memory_kernels.log 2>&1
tail -2 gpurun_out/r02_prof_full.log gpurun_out/r02_memory_kernels.log gpurun_out/r02_bench_under_ncu.err
