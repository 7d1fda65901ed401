# last GPU call of round 2: region graph after the counting-sort change, then as much of the full `-m gpu` suite as fits
mkdir -p gpurun_out
timeout 50 python tools/bench_agglomerate.py --z 64 --size 512 > gpurun_out/r02c_bench_agglomerate.json 2> gpurun_out/r02c_bench_agglomerate.err
cut -c1-900 gpurun_out/r02c_bench_agglomerate.json; tail -3 gpurun_out/r02c_bench_agglomerate.err
timeout 125 python -m pytest tests/test_segmentation_agglomerate.py tests -m gpu -x -q --durations=6 2>&1 | tail -22 > gpurun_out/r02c_gpu_tests_full.txt
tail -14 gpurun_out/r02c_gpu_tests_full.txt
