#!/bin/bash
# Final GPU session of a round: smoke, the whole -m gpu suite, the bench line, the ncu launch list
# of the bench command, a sweep over the other workloads, full-size ncu captures for traffic.json.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))"; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -n 3 --durations=10 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))"; tail -4 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? t=$((SECONDS-T0))"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$? t=$((SECONDS-T0))"
timeout 400 bash tools/sweep.sh c2c256 c2c1024 c2c2048 c2c8192 c2c32768 c2c1m r2c8192 c2r8192 r2c16384 c2c2d512 c2c2d2048 r2c2d1024 c2r2d1024 > gpurun_out/sweep.log 2>&1; echo "sweep t=$((SECONDS-T0))"; cat gpurun_out/sweep.log
for wl in r2c4096 c2r4096 conv8192 c2c2d1024 c2c64; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:fft -s 3 -c 2 -f -o gpurun_out/prof_${wl}_full \
      python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu --also '' > gpurun_out/ncu_$wl.log 2>&1; echo "ncu $wl rc=$? t=$((SECONDS-T0))"
done
