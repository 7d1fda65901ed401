# compute-sanitizer passes for a GPU box (race detection on the blend / halo kernels).  Small cases only: the
# sanitizer slows kernels ~50-100x.      gpurun --timeout 1200 -- 'bash tools/sanitize.sh > gpurun_out/r02_sanitizer.txt 2>&1'
export CFB_NO_AUTOTUNE=1   # one tile configuration per layer instead of ~50 timed candidates
run() { echo "=== $*"; "$@" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|error|Error|RACECHECK SUMMARY|hazard|smoke:" | tail -12; }
SEL='identity_nonaligned or test_edge_cases or nan_and_overflow or quantize_maskout_crop_golden or normalize_contrast_golden or slab_entry_point'
# memcheck: identity blend / extract / weight volume / normalise / range check / slab path / operators
run compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_operators.py -x -q -k "$SEL"
# memcheck on the tcgen05 network path: smoke (f16f8 default: TMEM-shift kernel, fused tail, convT) and the reference-literal --augment
run compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()"
run compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -k "test_unet3l_golden or (test_device_test_time_augmentation and True-None)"
# racecheck (shared-memory hazards) on the CUDA-core kernels with shared-memory staging; it does not model mbarrier / async-proxy
# ordering, so the tcgen05 kernels get synccheck instead
run compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_operators.py tests/test_gpu_parity.py -x -q -k "normalize_contrast_golden or quantize_maskout_crop_golden or identity_nonaligned"
run compute-sanitizer --tool synccheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()"
# initcheck: reads of uninitialised device memory on the identity path (accumulator memset, weight volume, halo add)
run compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -k "identity_nonaligned or slab_entry_point"
