#!/bin/bash
# round 2, GPU call H: four-step column kernel vs two sweeps; row-repeat variants of the fused 2D kernel; full suite
mkdir -p gpurun_out
FB="python tools/fastbench.py"
$FB --label default cols32768x4096 cols32768x2048 cols16384x4096 cols8192x8192 cols4096x16384 > gpurun_out/r02h_fast_cols.jsonl 2>&1
$FB --label two_kernel --flags 0x200000 cols32768x4096 cols16384x4096 cols8192x8192 cols4096x16384 >> gpurun_out/r02h_fast_cols.jsonl 2>&1
bash tools/ab_fast.sh tools/ab_r2h.txt > gpurun_out/r02h_ab.txt 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_pytest.log
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
$NCU -k regex:fft_longcols -o gpurun_out/r02h_longcols32768 python tools/fastbench.py --reps 1 cols32768x4096 > gpurun_out/r02h_ncu.log 2>&1
cat gpurun_out/r02h_fast_cols.jsonl | cut -c1-160; cat gpurun_out/r02h_ab.txt; tail -3 gpurun_out/r02h_pytest.log
