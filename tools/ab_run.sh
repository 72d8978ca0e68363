#!/bin/bash
# Runs on the GPU box: measures the prebuilt variants of tools/ab_build.sh (same config file).
# The product library is put back at the end.  Results: gpurun_out/ab.jsonl + a table on stdout.
cfg=${1:-tools/ab.txt}
out=gpurun_out/ab.jsonl
mkdir -p gpurun_out; : > $out
cp mufft_b200/libmufft_b200.so /tmp/libmufft_b200.product.so
while IFS='|' read -r LABEL EXTRA WLS CHECKS; do
  LABEL=$(echo $LABEL); [ -z "$LABEL" ] && continue
  case "$LABEL" in \#*) continue;; esac
  if [ "$LABEL" = base ]; then cp /tmp/libmufft_b200.product.so mufft_b200/libmufft_b200.so
  else
    [ -f ab_variants/$LABEL/libmufft_b200.so ] || { echo "$LABEL | not built"; continue; }
    cp ab_variants/$LABEL/libmufft_b200.so mufft_b200/libmufft_b200.so
  fi
  ok=""
  for n in $CHECKS; do ok="$ok $n:$(timeout 300 python tools/quick_check.py $n 2>&1 | tail -1 | cut -c1-40)"; done
  echo "== $LABEL [$EXTRA] check:$ok"
  for w in $WLS; do
    wl=${w%%:*}; units=""; [ "$w" != "$wl" ] && units="--units ${w##*:}"
    res=$(timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu --also '' $units 2>&1 | tail -1)
    echo "{\"label\": \"$LABEL\", \"extra\": \"$EXTRA\", \"workload\": \"$wl\", \"check\": \"$ok\", \"bench\": $res}" >> $out
    python - "$wl" <<PY
import json,sys
try:
    r=json.loads('''$res''')
    print("   %-10s %.3f ms  %.0f GB/s  frac %.3f  (min step %.3f ms)" % (sys.argv[1], r["ms_per_step"], r["roofline"]["achieved"], r["roofline"]["frac"], r["roofline"]["min_step_ms"]))
except Exception as e:
    print("   %-10s bench failed:" % sys.argv[1], '''$res'''[-300:])
PY
  done
done < $cfg
cp /tmp/libmufft_b200.product.so mufft_b200/libmufft_b200.so
