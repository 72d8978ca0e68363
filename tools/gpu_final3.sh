#!/bin/bash
# Validation after the partitioned-convolution load modes (all row kernels were recompiled).
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -n 3 --durations=6 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/pytest.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | head
timeout 300 python bench.py --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$? t=$((SECONDS-T0))"; cut -c1-200 gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
timeout 300 bash tools/sweep.sh conv8192_part8 conv8192 slab2d8k slab2d8k_r2c > gpurun_out/sweep3.log 2>&1; echo "sweep t=$((SECONDS-T0))"; cat gpurun_out/sweep3.log
