#!/bin/bash
# round 2, GPU call D: split four-step kernel vs the one-transform-at-a-time variants; where a small call's microseconds go
mkdir -p gpurun_out
bash tools/ab_fast.sh tools/ab_r2d.txt > gpurun_out/r02d_ab.txt 2>&1
FB="python tools/fastbench.py"
$FB --label default c2c32768:16384 c2c65536:8192 r2c65536 c2r65536 > gpurun_out/r02d_fast_default.jsonl 2>&1
$FB --label four_step_16k --flags 0x800000 c2c16384 r2c32768 c2r32768 > gpurun_out/r02d_fast_fourstep16k.jsonl 2>&1
tools/_build/latency_probe 20000 > gpurun_out/r02d_latency.jsonl 2>&1
tools/_build/latency_probe 20000 spin > gpurun_out/r02d_latency_spin.jsonl 2>&1
python -m pytest tests/test_batch_gpu.py tests/test_slab_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 -k "large_n or side_by_side or slab_plan" > gpurun_out/r02d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_pytest.log
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
$NCU -k regex:fft_long -o gpurun_out/r02d_long32768 python tools/fastbench.py --reps 1 c2c32768 > gpurun_out/r02d_ncu.log 2>&1
cat gpurun_out/r02d_ab.txt; cat gpurun_out/r02d_latency.jsonl; tail -3 gpurun_out/r02d_pytest.log
