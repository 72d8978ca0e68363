timeout 600 python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "2d" 2>&1 | tail -2
for f in 1; do
  MUFFT_BENCH_FUSED_2D=$f timeout 300 python bench.py --workload c2c2d1024 --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('fused=$f c2c2d1024 %.3f ms %.0f GB/s frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['achieved'], r['roofline']['frac'], r['gpu_launches']))"
done
printf 'c2c2d1024 256 1\n' > /tmp/p.txt; bash tools/gpu_prof.sh /tmp/p.txt
