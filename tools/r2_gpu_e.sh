#!/bin/bash
# round 2, GPU call E: fused 2D split kernel vs the ticket kernel; graph-replay latency path; new GPU tests
mkdir -p gpurun_out
bash tools/ab_fast.sh tools/ab_r2e.txt > gpurun_out/r02e_ab.txt 2>&1
tools/_build/latency_probe 20000 > gpurun_out/r02e_latency.jsonl 2>&1
python -m pytest tests/test_batch_gpu.py tests/test_parity_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "graph or c2c_1d or conv or c2c_2d" > gpurun_out/r02e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_pytest.log
cp ab_variants/f2d_split/libmufft_b200.so /tmp/split.so
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
MUFFT_B200_LIB=/tmp/split.so $NCU -k regex:fft2d_split -o gpurun_out/r02e_c2c2d1024_split python tools/fastbench.py --reps 1 c2c2d1024 > gpurun_out/r02e_ncu.log 2>&1
cat gpurun_out/r02e_ab.txt; cat gpurun_out/r02e_latency.jsonl; tail -3 gpurun_out/r02e_pytest.log
