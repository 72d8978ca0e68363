#!/bin/bash
# round 2, GPU call C: full suite on the refactored tree, long-kernel variants, fused C2R after the fence change
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/r02c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest.log
FB="python tools/fastbench.py"
bash tools/ab_fast.sh tools/ab_r2c.txt > gpurun_out/r02c_ab.txt 2>&1
$FB --label default c2c32768:16384 c2c65536:8192 c2r2d1024 c2r2d512 c2r2d2048 r2c2d1024 c2c2d1024 r2c65536 c2r65536 > gpurun_out/r02c_fast_default.jsonl 2>&1
$FB --label four_step_16k --flags 0x800000 c2c16384 > gpurun_out/r02c_fast_fourstep16k.jsonl 2>&1
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
$NCU -k regex:fft_long_kernel -o gpurun_out/r02c_long32768 python tools/fastbench.py --reps 1 c2c32768 > gpurun_out/r02c_ncu.log 2>&1
tail -4 gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_ab.txt
