#!/bin/bash
# The records as the driver takes them (run under gpurun on one B200; outputs under gpurun_out/r02l_*, copied to profiles/) - full suite with -x, smoke, bench of record,
# reference arm, ncu launch list of the bench command, ncu --set full of the headline kernel
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02l_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_smoke.log
python bench.py --gpus 1 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "rc=$?" >> gpurun_out/r02l_bench.err
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02l_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --also '' > gpurun_out/r02l_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f -k regex:fft_rows_kernel -o gpurun_out/r02l_c2c4096 python tools/fastbench.py --reps 1 c2c4096:262144 > gpurun_out/r02l_ncu.log 2>&1
python tools/fastbench.py --label final c2c4096:262144 r2c4096:262144 c2r4096:262144 conv8192:65536 conv8192_fused:65536 c2c2d1024 r2c2d1024 c2r2d1024 c2c16384 c2c32768:16384 c2c65536:8192 c2c2d256 c2c2d512 c2c2d2048 > gpurun_out/r02l_fast_final.jsonl 2>&1
tail -3 gpurun_out/r02l_pytest.log; tail -2 gpurun_out/r02l_smoke.log; wc -c gpurun_out/r02l_bench.json; cut -c1-600 gpurun_out/r02l_bench.json; cut -c1-300 gpurun_out/r02l_bench_reference.json
