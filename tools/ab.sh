#!/bin/bash
# A/B measurements on the GPU box: every line of the config file ($1) is
#   <label> | <EXTRA nvcc flags> | <workload[:units] ...> | <lengths for tools/quick_check.py>
# The library is rebuilt with those flags, spot-checked against the oracle, and each workload is
# timed with bench.py (device-resident arm only).  Results: gpurun_out/ab.jsonl + a table on stdout.
cfg=${1:-tools/ab.txt}
out=gpurun_out/ab.jsonl
mkdir -p gpurun_out; : > $out
while IFS='|' read -r LABEL EXTRA WLS CHECKS; do
  LABEL=$(echo $LABEL); [ -z "$LABEL" ] && continue
  case "$LABEL" in \#*) continue;; esac
  make -s -C mufft_b200/csrc clean >/dev/null
  if ! make -s -C mufft_b200/csrc -j 32 EXTRA="$EXTRA" > gpurun_out/ab_build.log 2>&1; then
    echo "$LABEL | build failed"; tail -5 gpurun_out/ab_build.log; continue
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
make -s -C mufft_b200/csrc clean >/dev/null
