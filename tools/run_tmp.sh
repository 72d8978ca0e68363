timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
for f in 0 1; do
  MUFFT_BENCH_FUSED_2D=$f timeout 300 python bench.py --workload c2c2d1024 --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('fused=$f c2c2d1024 %.3f ms %.0f GB/s frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['achieved'], r['roofline']['frac'], r['gpu_launches']))"
done
printf 'c2c2d1024 256 1\n' > /tmp/p.txt; bash tools/gpu_prof.sh /tmp/p.txt
