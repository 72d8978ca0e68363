#!/bin/bash
# round 2, GPU call B: the refactored tree on hardware - full suite, smoke, fastbench of every path incl. the
# fallback flags, A/B variants, ncu captures of the new kernels, bench of record
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -x > gpurun_out/r02b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02b_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_smoke.log
FB="python tools/fastbench.py"
$FB --label default c2c4096 r2c4096 c2r4096 c2c16384 c2c32768 c2c65536 r2c65536 c2r65536 c2c131072 conv8192 conv8192_fused c2c2d256 c2c2d512 c2c2d1024 c2c2d2048 r2c2d1024 c2r2d1024 r2c2d512 c2r2d512 r2c2d2048 c2r2d2048 r2c2d4096x1024 c2r2d4096x1024 > gpurun_out/r02b_fast_default.jsonl 2>&1
$FB --label two_kernel --flags 0x200000 c2c16384 c2c32768 c2c65536 > gpurun_out/r02b_fast_two_kernel.jsonl 2>&1
$FB --label four_step_16k --flags 0x800000 c2c16384 r2c32768 c2r32768 > gpurun_out/r02b_fast_fourstep16k.jsonl 2>&1
$FB --label separate_2d --flags 0x100000 c2c2d1024 r2c2d1024 c2r2d1024 c2r2d512 c2r2d2048 > gpurun_out/r02b_fast_separate.jsonl 2>&1
bash tools/ab_fast.sh tools/ab_r2b.txt > gpurun_out/r02b_ab.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f"
$NCU -k regex:fft_long_kernel -o gpurun_out/r02b_long32768 python tools/fastbench.py --reps 1 c2c32768 > gpurun_out/r02b_ncu.log 2>&1
$NCU -k regex:fft2d_fused_kernel -o gpurun_out/r02b_c2c2d1024 python tools/fastbench.py --reps 1 c2c2d1024 >> gpurun_out/r02b_ncu.log 2>&1
$NCU -k regex:fft2d_fused_kernel -o gpurun_out/r02b_c2r2d1024 python tools/fastbench.py --reps 1 c2r2d1024 >> gpurun_out/r02b_ncu.log 2>&1
python bench.py --gpus 1 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "rc=$?" >> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_pytest.log; tail -2 gpurun_out/r02b_smoke.log; cat gpurun_out/r02b_fast_default.jsonl | cut -c1-150
