# development: timing sensitivity of the TMEM-shift conv kernel (results become wrong).  Needs the instrumented build:
#   python -m chunkflow_b200.build --variant ablate -DCFB_TS_ABLATE
# bits of CFB_ABLATE: 1 no global stores, 2 no epilogue TMEM reads, 4 no loader copies, 8 no shifts, 16 no MMAs, 32 no TMA loads
for m in 0 1 3 4 8 16 24 32 36 7 63; do
  CFB_NATIVE_LIB=$PWD/chunkflow_b200/_native/libchunkflow_b200_ablate.so CFB_ABLATE=$m timeout 300 python bench.py --workload 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_chunk']
print('ablate=%3s' % '$m', 'step %.1f' % d['ms_per_step'], ' '.join('%s=%.2f' % (n, k[n]) for n in ['enc0.2','enc1.0','enc1.2','enc2.0','dec1.0','dec1.2','dec0.0','dec0.2+head+blend']))"
done
