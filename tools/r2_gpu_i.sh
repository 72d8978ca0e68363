#!/bin/bash
# round 2, GPU call I: staging pipeline of the host-pointer path (slots x chunk size), raw bus rate beside it
mkdir -p gpurun_out
bash tools/e2e_sweep.sh ab_variants/stage_debug/libmufft_b200.so > gpurun_out/r02i_e2e_sweep.txt 2>&1
python tools/fastbench.py --reps 3 e2e4096:262144 > gpurun_out/r02i_e2e_product.jsonl 2>&1
cat gpurun_out/r02i_e2e_sweep.txt | cut -c1-230; cat gpurun_out/r02i_e2e_product.jsonl | cut -c1-300
