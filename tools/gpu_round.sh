#!/bin/bash
# One GPU session: smoke, parity tests, bench, ncu launch list and a full capture of the row kernel.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --also r2c4096,conv8192,c2c2d1024 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_rows -s 2 -c 2 -f -o gpurun_out/prof_rows \
    python bench.py --steps 1 --warmup 3 --units 65536 --no-e2e --no-cpu --also '' > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
