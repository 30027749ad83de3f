#!/bin/bash
# like exp_run.sh, BG2 only, two interleaved rounds (box-to-box and run-to-run noise is a few percent: compare within one call)
OUT=$1; shift
: > "$OUT"
for round in 1 2; do
for v in "$@"; do
  if [ "$v" = base ]; then unset NLDPC_LIB_PATH; else export NLDPC_LIB_PATH=$PWD/build/exp/$v.so; fi
  for mode in packed list; do
    echo -n "$v " >> "$OUT"; python tools/prof_decode.py $mode 65536 nr_bg2_set0 >> "$OUT" 2>&1
  done
done
done
cat "$OUT"
