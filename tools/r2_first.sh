#!/bin/bash
# First GPU call of round 2: measures every path that round 1 built and validated on the CPU emulation
# but could not time (profiles/README.md section 6).  No torch import, about a minute of box time:
#   bash tools/ab_build.sh tools/ab_r2.txt          # here, before the call (builds ab_variants/*)
#   gpurun --timeout 300 -- 'bash tools/r2_first.sh'
# Results: gpurun_out/probe_optin.jsonl, gpurun_out/fastbench.jsonl, gpurun_out/r2_first.log
mkdir -p gpurun_out
{
  for what in early real2d managed latency cluster; do
    echo "=== probe $what"
    timeout 120 python tools/probe_optin.py $what 2>&1 | cut -c1-400
  done
  echo "=== experimental GPU tests that need no torch-free probe: slab R2C p2p (1 GPU)"
  MUFFT_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q 2>&1 | tail -15
  echo "=== A/B builds"
  bash tools/ab_fast.sh tools/ab_r2.txt
  echo "=== fused 2D at three CTAs per SM needs a flavour without shared-memory tables"
  if [ -f ab_variants/f2d_mb3/libmufft_b200.so ]; then
    cp mufft_b200/libmufft_b200.so /tmp/prod.so; cp ab_variants/f2d_mb3/libmufft_b200.so mufft_b200/libmufft_b200.so
    for mode in 6 7 10; do MUFFT_CUDA_FUSED_2D=$mode timeout 60 python tools/fastbench.py --label f2d_mb3_mode$mode c2c2d1024 c2c2d512 2>&1 | cut -c1-300; done
    cp /tmp/prod.so mufft_b200/libmufft_b200.so
  fi
} 2>&1 | tee gpurun_out/r2_first.log
