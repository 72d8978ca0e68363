#!/bin/bash
# One GPU session: smoke, parity tests, bench, ncu launch list of the bench command.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
T0=$SECONDS
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$? t=$((SECONDS-T0))"
