#!/bin/bash
# Builds A/B variants of the library HERE (no GPU needed): every line of the config file ($1) is
#   <label> | <EXTRA nvcc flags> | <workload[:units] ...> | <objects to rebuild, e.g. "f2d_* long_*"; empty = everything>
# -> ab_variants/<label>/libmufft_b200.so.  tools/ab_run.sh then measures them on the GPU box
# without spending GPU time on compiling.
cfg=${1:-tools/ab_example.txt}
mkdir -p ab_variants
while IFS='|' read -r LABEL EXTRA WLS CHECKS; do
  LABEL=$(echo $LABEL); [ -z "$LABEL" ] && continue
  case "$LABEL" in \#*) continue;; esac
  [ "$LABEL" = base ] && continue  # the product library itself
  mkdir -p ab_variants/$LABEL
  ONLY=$(echo $CHECKS)
  if [ -n "$ONLY" ] && [ -d mufft_b200/csrc/build ]; then
    # start from the product's objects; only the named ones are compiled with the variant's flags
    cp -r mufft_b200/csrc/build mufft_b200/csrc/build_ab_$LABEL
    for g in $ONLY; do rm -f mufft_b200/csrc/build_ab_$LABEL/$g.o; done
  fi
  if make -s -C mufft_b200/csrc -j 16 BUILD=build_ab_$LABEL OUT=../../ab_variants/$LABEL/libmufft_b200.so EXTRA="$EXTRA" > ab_variants/$LABEL/build.log 2>&1; then
    echo "$LABEL built"
  else
    echo "$LABEL BUILD FAILED"; tail -5 ab_variants/$LABEL/build.log
  fi
  rm -rf mufft_b200/csrc/build_ab_$LABEL
done < $cfg
