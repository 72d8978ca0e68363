#!/bin/bash
# Re-validation after the split / grouping / staging changes: whole -m gpu suite, bench line, affected workloads.
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -n 3 --durations=6 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/pytest.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest.log | head
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 400 bash tools/sweep.sh c2c16384 c2c32768 c2c1m r2c2d1024 c2r2d1024 slab2d8k slab2d8k_r2c > gpurun_out/sweep2.log 2>&1; echo "sweep t=$((SECONDS-T0))"; cat gpurun_out/sweep2.log
