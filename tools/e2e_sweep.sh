#!/bin/bash
# Staging-pipeline sweep of the host-pointer path (tuning runs only): a library built with -DMUFFT_CUDA_DEBUG reads the
# number of slots and the chunk size from the environment.  usage: tools/e2e_sweep.sh <path to the debug build>
lib=${1:-ab_variants/stage_debug/libmufft_b200.so}
for slots in 3 4 6; do
  for mb in 32 64 128 256; do
    echo -n "slots=$slots chunk_mb=$mb  "
    MUFFT_B200_LIB=$lib MUFFT_CUDA_STAGE_SLOTS=$slots MUFFT_CUDA_STAGE_CHUNK_MB=$mb python tools/fastbench.py --reps 3 e2e4096:131072 2>&1 | tail -1 | cut -c1-220
  done
done
