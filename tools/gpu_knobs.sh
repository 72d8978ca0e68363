#!/bin/bash
# Runtime-knob measurements (no rebuild): 2D intermediate group size, two-kernel split, fused vs separate 2D.
mkdir -p gpurun_out; : > gpurun_out/knobs.log
run() { # label, env..., -- workload
  label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  res=$(env "${envs[@]}" timeout 200 python bench.py --workload $1 --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>&1 | tail -1)
  echo "$res" | python -c "
import json,sys
try:
    r=json.loads(sys.stdin.read()); print('%-34s %-12s %8.3f ms  frac %.3f  launches/step %.0f' % ('$label', '$1', r['ms_per_step'], r['roofline']['frac'], r['gpu_launches']/r['steps']))
except Exception as e: print('$label $1 failed', e)
" | tee -a gpurun_out/knobs.log
}
for mb in 48 96 192 4096; do
  run "mid_group=${mb}MiB" MUFFT_CUDA_MID_GROUP_MB=$mb -- r2c2d1024
  run "mid_group=${mb}MiB" MUFFT_CUDA_MID_GROUP_MB=$mb -- c2r2d1024
done
for n2 in 256 512 1024; do
  run "split_n2=$n2" MUFFT_CUDA_SPLIT_N2=$n2 -- c2c32768
  run "split_n2=$n2" MUFFT_CUDA_SPLIT_N2=$n2 -- c2c1m
done
for mb in 48 4096; do
  run "separate kernels, mid_group=${mb}MiB" MUFFT_BENCH_FUSED_2D=0 MUFFT_CUDA_MID_GROUP_MB=$mb -- c2c2d2048
  run "separate kernels, mid_group=${mb}MiB" MUFFT_BENCH_FUSED_2D=0 MUFFT_CUDA_MID_GROUP_MB=$mb -- c2c2d512
  run "separate kernels, mid_group=${mb}MiB" MUFFT_BENCH_FUSED_2D=0 MUFFT_CUDA_MID_GROUP_MB=$mb -- c2c2d1024
done
run "prefetch kernel" MUFFT_CUDA_PREFETCH=1 -- c2c2048
run "prefetch kernel" MUFFT_CUDA_PREFETCH=1 -- c2c8192
run "prefetch kernel" MUFFT_CUDA_PREFETCH=1 -- conv8192
