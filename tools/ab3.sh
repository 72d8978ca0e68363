#!/bin/bash
make -s -C mufft_b200/csrc -j32 >/dev/null 2>&1
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_slab_gpu.py -m gpu -x -q -k "large_n or slab" 2>&1 | tail -3
for mb in 8 16 32 64 1024; do for w in c2c32768 c2c1m; do
MUFFT_CUDA_L2_CHUNK_MB=$mb python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | grep "^{" | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('chunk=$mb MiB $w %.3f ms frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['frac'], r['gpu_launches']))"
done; done
for mb in 16 32 1024; do
MUFFT_CUDA_L2_CHUNK_MB=$mb python bench.py --workload slab2d8k --steps 10 --warmup 3 --also '' 2>&1 | grep "^{" | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('chunk=$mb MiB slab2d8k(1 gpu) %.3f ms frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['frac'], r['gpu_launches']))"
MUFFT_CUDA_L2_CHUNK_MB=$mb python bench.py --workload slab2d --steps 3 --warmup 3 --also '' 2>&1 | grep "^{" | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('chunk=$mb MiB slab2d(1 gpu) %.3f ms frac %.3f launches %d' % (r['ms_per_step'], r['roofline']['frac'], r['gpu_launches']))"
done
