for m in 0 1 3 4 8 16 24 32 36 7 63; do
  CFB_ABLATE=$m timeout 300 python bench.py --workload 256 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_chunk']
print('ablate=%3s' % '$m', 'step %.1f' % d['ms_per_step'], ' '.join('%s=%.2f' % (n, k[n]) for n in ['enc0.2','enc1.0','enc1.2','enc2.0','dec1.0','dec1.2','dec0.0','dec0.2+head+blend']))"
done
