#!/bin/bash
# ncu --set full captures of the dominant kernel for a list of "workload units [skip]" lines (stdin or $1).
# Output: gpurun_out/prof_<workload>.ncu-rep
mkdir -p gpurun_out
while read -r WL UNITS SKIP; do
  [ -z "$WL" ] && continue
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft -s ${SKIP:-4} -c 2 -f -o gpurun_out/prof_$WL \
    python bench.py --workload $WL --steps 1 --warmup 3 --units $UNITS --no-e2e --no-cpu --also '' > gpurun_out/ncu_$WL.log 2>&1
  echo "$WL rc=$?"
done < ${1:-/dev/stdin}
