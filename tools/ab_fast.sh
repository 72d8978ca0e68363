#!/bin/bash
# Runs on the GPU box: measures prebuilt variants (tools/ab_build.sh, same config file) with
# tools/fastbench.py - no torch import, a few seconds per variant.  Config lines:
#   <label> | <EXTRA nvcc flags> | <fastbench workloads> | <ignored>
# The product library is put back at the end.  Results: gpurun_out/fastbench.jsonl (label = variant).
cfg=${1:-tools/ab_example.txt}
mkdir -p gpurun_out
cp mufft_b200/libmufft_b200.so /tmp/libmufft_b200.product.so
while IFS='|' read -r LABEL EXTRA WLS CHECKS; do
  LABEL=$(echo $LABEL); [ -z "$LABEL" ] && continue
  case "$LABEL" in \#*) continue;; esac
  if [ "$LABEL" = base ]; then cp /tmp/libmufft_b200.product.so mufft_b200/libmufft_b200.so
  else
    [ -f ab_variants/$LABEL/libmufft_b200.so ] || { echo "$LABEL | not built"; continue; }
    cp ab_variants/$LABEL/libmufft_b200.so mufft_b200/libmufft_b200.so
  fi
  echo "== $LABEL [$EXTRA]"
  timeout 120 python tools/fastbench.py --label "$LABEL" $WLS 2>&1 | python -c '
import sys, json
for line in sys.stdin:
    try:
        r = json.loads(line)
        print("   %-16s %8.3f ms  %7.0f GB/s  frac %.3f  %s" % (r["workload"], r["ms"], r["gbs"], r["frac"], "ok" if r["ok"] else "PARITY FAILED %.2e" % r["rel_l2"]))
    except Exception:
        print("   " + line.rstrip())
'
done < $cfg
cp /tmp/libmufft_b200.product.so mufft_b200/libmufft_b200.so
