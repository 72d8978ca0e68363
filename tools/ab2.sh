#!/bin/bash
# unfused 2D vs fused
make -s -C mufft_b200/csrc -j32 >/dev/null 2>&1
for f in 1 0; do for w in c2c2d1024 c2c2d512 c2c2d2048; do
MUFFT_BENCH_FUSED_2D=$f python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('fused=$f $w %.3f ms frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['frac'], r['gpu_launches']))"
done; done
