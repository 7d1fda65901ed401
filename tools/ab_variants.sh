# development: A/B instrumented / alternative builds of the native library on the 256x512x512 workload (99 patches)
#   tools/ab_variants.sh <precision> <variant> [<variant> ...]     variant "product" = the in-tree product library,
#   anything else = chunkflow_b200/_native/libchunkflow_b200_<variant>.so (python -m chunkflow_b200.build --variant <name> -D...)
prec=$1; shift
for v in "$@"; do
  lib=chunkflow_b200/_native/libchunkflow_b200_$v.so
  [ "$v" = product ] && lib=chunkflow_b200/_native/libchunkflow_b200.so
  CFB_NATIVE_LIB=$PWD/$lib CFB_TS_TRACE_PRINT=1 timeout 300 python bench.py --workload 256 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-pageable --precision $prec 2>gpurun_out/ab_${prec}_$v.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_chunk']
print('$prec $v', 'step %.1f' % d['ms_per_step'], ' '.join('%s=%.2f' % (n, k[n]) for n in k))"
  grep cfb-trace gpurun_out/ab_${prec}_$v.err | awk '!seen[$2 $3]++' | cut -c1-420
done
