#!/bin/bash
# The records as the driver takes them (run under gpurun on one B200; outputs under gpurun_out/<prefix>_*, copied to
# profiles/): full suite with -x, smoke, bench of record, reference arm, ncu launch list of the bench command, ncu --set
# full of the headline kernel, of the TMA 2D kernel and of the prefetching row kernel, every workload once more.
P=${1:-r02fin}
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/${P}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${P}_pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${P}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${P}_smoke.log
python bench.py --gpus 1 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "rc=$?" >> gpurun_out/${P}_bench.err
python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --also '' > gpurun_out/${P}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 2 -c 1 -f -k regex:fft_rows_kernel -o gpurun_out/${P}_c2c4096 python tools/fastbench.py --reps 1 c2c4096:262144 > gpurun_out/${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 1 -c 1 -f -k regex:fft2d_tma -o gpurun_out/${P}_c2c2d1024_tma python tools/fastbench.py --reps 1 c2c2d1024 >> gpurun_out/${P}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 1 -c 1 -f -k regex:fft_rows_prefetch -o gpurun_out/${P}_c2c16384_prefetch python tools/fastbench.py --reps 1 c2c16384 >> gpurun_out/${P}_ncu.log 2>&1
python tools/fastbench.py --label final c2c4096:262144 r2c4096:262144 c2r4096:262144 c2c8192 c2c16384 conv8192:65536 conv8192_fused:65536 c2c2d1024 r2c2d1024 c2r2d1024 c2c2d512x1024 c2c32768:16384 c2c65536:8192 cols32768x4096 r2c8192 c2r8192 r2c16384 c2r16384 c2c2d256 c2c2d512 c2c2d2048 c2c64 c2c1024 c2c2048 > gpurun_out/${P}_fast_final.jsonl 2>&1
tail -3 gpurun_out/${P}_pytest.log; tail -2 gpurun_out/${P}_smoke.log; wc -c gpurun_out/${P}_bench.json; cut -c1-600 gpurun_out/${P}_bench.json; cut -c1-300 gpurun_out/${P}_bench_reference.json
