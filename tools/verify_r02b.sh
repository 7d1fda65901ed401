# round-2 (second session) GPU verification of the agglomerate operator and the operator changes; run under gpurun, one GPU.
# (the network / blend kernels are untouched in this session: their tests run in the driver's round-end `pytest -m gpu`)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_segmentation_agglomerate.py tests/test_segmentation.py tests/test_gpu_operators.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02b_gpu_tests.txt
tail -3 gpurun_out/r02b_gpu_tests.txt
timeout 100 python tools/bench_agglomerate.py > gpurun_out/r02b_bench_agglomerate.json 2> gpurun_out/r02b_bench_agglomerate.err
timeout 60 python tools/bench_operators.py > gpurun_out/r02b_bench_operators.json 2> gpurun_out/r02b_bench_operators.err
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_segmentation_agglomerate.py tests/test_gpu_operators.py -m gpu -q -x -k "watershed_and_region or plugin_and_cli or random_vs_oracle" > gpurun_out/r02b_sanitizer_agglomerate.txt 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r02b_sanitizer_agglomerate.txt
tail -4 gpurun_out/r02b_sanitizer_agglomerate.txt
cut -c1-900 gpurun_out/r02b_bench_agglomerate.json; tail -3 gpurun_out/r02b_bench_agglomerate.err
python -c "
import json
d=json.load(open('gpurun_out/r02b_bench_operators.json'))
print({k:(round(v['ms'],3),round(v['frac_of_hbm_peak'],3)) for k,v in d['operators'].items()})"
