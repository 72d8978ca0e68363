#!/bin/bash
# Measures alternative engine shapes on the GPU box: for every line "N R r0 r1 r2 r3 CTA WORKLOAD [UNITS]"
# in the config file ($1) the library is rebuilt with that override, checked against the oracle on
# that length, and timed with bench.py (device-resident arm only).  Results: gpurun_out/tune.jsonl
cfg=${1:-tools/tune_shapes.txt}
out=gpurun_out/tune.jsonl
mkdir -p gpurun_out; : > $out
while read -r N R r0 r1 r2 r3 CTA WL UNITS; do
  [ -z "$N" ] && continue
  case "$N" in \#*) continue;; esac
  EXTRA="-DMUFFT_SHAPE_N=$N -DMUFFT_SHAPE_R=$R -DMUFFT_SHAPE_R0=$r0 -DMUFFT_SHAPE_R1=$r1 -DMUFFT_SHAPE_R2=$r2 -DMUFFT_SHAPE_R3=$r3 -DMUFFT_ROWS_CTA_THREADS=$CTA $XFLAGS"
  make -s -C mufft_b200/csrc clean >/dev/null
  if ! make -s -C mufft_b200/csrc -j 16 EXTRA="$EXTRA" > gpurun_out/tune_build.log 2>&1; then
    echo "{\"shape\": \"$N $R $r0 $r1 $r2 $r3 $CTA\", \"error\": \"build failed\"}" >> $out; continue
  fi
  ok=$(timeout 300 python tools/quick_check.py $N 2>&1 | tail -1)
  res0=$(MUFFT_CUDA_PREFETCH=0 timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-e2e --no-cpu --also '' ${UNITS:+--units $UNITS} 2>&1 | tail -1)
  res=$(timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-e2e --no-cpu --also '' ${UNITS:+--units $UNITS} 2>&1 | tail -1)
  echo "{\"shape\": \"$N $R $r0 $r1 $r2 $r3 cta=$CTA $XFLAGS\", \"workload\": \"$WL\", \"check\": \"$ok\", \"bench\": $res, \"bench_noprefetch\": $res0}" >> $out
  python - "$N $R $r0,$r1,$r2,$r3 cta=$CTA $WL" "$ok" <<PY
import json,sys
try:
    r=json.loads('''$res'''); r0=json.loads('''$res0''')
    print(sys.argv[1], "|", sys.argv[2], "| prefetch: %.3f ms %.0f GB/s frac %.3f | direct: %.3f ms frac %.3f" % (r["ms_per_step"], r["roofline"]["achieved"], r["roofline"]["frac"], r0["ms_per_step"], r0["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "| bench failed:", '''$res'''[-300:], '''$res0'''[-200:])
PY
done < $cfg
make -s -C mufft_b200/csrc clean >/dev/null
