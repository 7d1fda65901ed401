# development: run the 256^3 workload with instrumented builds of the native library (see chunkflow_b200/build.py --variant)
for v in "$@"; do
  lib=chunkflow_b200/_native/libchunkflow_b200_$v.so
  [ "$v" = product ] && lib=chunkflow_b200/_native/libchunkflow_b200.so
  CFB_NATIVE_LIB=$PWD/$lib CFB_TS_TRACE_PRINT=1 timeout 300 python bench.py --workload 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>gpurun_out/trace_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_chunk']
print('$v', 'step %.1f' % d['ms_per_step'], ' '.join('%s=%.2f' % (n, k[n]) for n in ['enc0.2','enc1.0','enc1.2','enc2.0','dec1.0','dec1.2','dec0.0','dec0.2+head+blend']))"
  grep cfb-trace gpurun_out/trace_$v.err | awk '!seen[$2 $3]++' | cut -c1-330
done
