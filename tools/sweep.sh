#!/bin/bash
# Device-resident throughput of every bench workload (one JSON line each) -> gpurun_out/sweep.jsonl
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for wl in ${@:-c2c64 c2c256 c2c1024 c2c2048 c2c4096 c2c8192 c2c32768 c2c1m r2c2048 r2c4096 r2c8192 r2c16384 c2r2048 c2r4096 c2r8192 c2r16384 conv8192 conv8192_fused c2c2d512 c2c2d1024 c2c2d2048 r2c2d1024 c2r2d1024 slab2d8k slab2d8k_r2c}; do
  res=$(timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --also '' 2>gpurun_out/sweep_err.log | tail -1)
  echo "$res" >> gpurun_out/sweep.jsonl
  python - "$wl" <<PY
import json,sys
try:
    r=json.loads('''$res''')
    print("%-16s %9.3f ms/step %8.0f GFLOP/s %7.0f GB/s  frac %.3f  launches/step %.1f" % (sys.argv[1], r["ms_per_step"], r["value"], r["roofline"]["achieved"], r["roofline"]["frac"], r["gpu_launches"] / r["steps"]))
except Exception as e:
    print("%-16s failed: %s" % (sys.argv[1], '''$res'''[-300:]))
PY
done
