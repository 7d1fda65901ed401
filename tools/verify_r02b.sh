# round-2 (second session) GPU verification of the agglomerate operator and the operator MLP change; run under gpurun, one GPU
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02b_gpu_tests.txt
tail -3 gpurun_out/r02b_gpu_tests.txt
timeout 120 python tools/bench_agglomerate.py > gpurun_out/r02b_bench_agglomerate.json 2> gpurun_out/r02b_bench_agglomerate.err
timeout 120 python tools/bench_operators.py > gpurun_out/r02b_bench_operators.json 2> gpurun_out/r02b_bench_operators.err
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_agglomeration.py -m gpu -q -x -k "watershed_and_region or plugin_and_cli" > gpurun_out/r02b_sanitizer_agglomerate.txt 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r02b_sanitizer_agglomerate.txt
tail -4 gpurun_out/r02b_sanitizer_agglomerate.txt
cat gpurun_out/r02b_bench_agglomerate.json | cut -c1-900
cat gpurun_out/r02b_bench_operators.json | cut -c1-1500
