# compute-sanitizer passes for a GPU box (race detection on the blend / halo kernels; not run in
# round 1 -- GPU minutes went to parity, ncu and the bench).  Small cases only: the sanitizer slows kernels ~50-100x.
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
set -x
export CFB_NO_AUTOTUNE=1   # one tile configuration per layer instead of ~50 timed candidates
SEL='identity_nonaligned or test_edge_cases or quantize_maskout_crop_golden or normalize_contrast_golden'
compute-sanitizer --tool memcheck  --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_operators.py -x -q -k "$SEL" 2>&1 | tail -15
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_operators.py -x -q -k "normalize_contrast_golden or quantize_maskout_crop_golden" 2>&1 | tail -15
# the tcgen05 kernels: memcheck on one small network call (racecheck does not model mbarrier / async-proxy ordering)
compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
