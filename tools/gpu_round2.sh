#!/bin/bash
# GPU session 2: new tests (reference consumers, fused convolution), A/B of prebuilt variants,
# bench line, full ncu captures of the headline kernel (full size) and the fused convolution.
mkdir -p gpurun_out
T0=$SECONDS
timeout 900 python -m pytest tests/test_consumers.py tests/test_batch_gpu.py tests/test_slab_gpu.py -m gpu -q -k "consumers or reference or conv or real" --durations=8 > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -12 gpurun_out/pytest2.log
timeout 900 bash tools/ab_run.sh tools/ab.txt 2>&1 | tee gpurun_out/ab.log; echo "ab t=$((SECONDS-T0))"
timeout 600 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$? t=$((SECONDS-T0))"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_rows -s 3 -c 1 -f -o gpurun_out/prof_c2c4096_full \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --also '' > gpurun_out/ncu_c2c4096.log 2>&1; echo "ncu c2c rc=$? t=$((SECONDS-T0))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_conv -s 3 -c 1 -f -o gpurun_out/prof_conv8192_fused \
    python bench.py --workload conv8192_fused --steps 1 --warmup 3 --no-e2e --no-cpu --also '' > gpurun_out/ncu_convf.log 2>&1; echo "ncu convf rc=$? t=$((SECONDS-T0))"
