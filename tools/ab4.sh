#!/bin/bash
make -s -C mufft_b200/csrc -j32 >/dev/null 2>&1
timeout 400 python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "2d" 2>&1 | tail -3
for w in c2c2d1024 c2c2d512 c2c2d2048; do
timeout 120 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | grep "^{" | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('$w %.3f ms frac %.3f' % (r['ms_per_step'], r['roofline']['frac']))"
done
