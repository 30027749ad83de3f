#!/bin/bash
# Experiment helper for the GPU box: time the decode kernels of several library variants (tools/build_variant.sh) back to back.
#   tools/exp_run.sh OUT.log variant1 variant2 ...      ("base" = the in-tree library)
OUT=$1; shift
: > "$OUT"
for v in "$@"; do
  if [ "$v" = base ]; then unset NLDPC_LIB_PATH; else export NLDPC_LIB_PATH=$PWD/build/exp/$v.so; fi
  for mode in packed list; do
    for code in nr_bg2_set0 wimax_n576_r34; do
      echo -n "$v " >> "$OUT"; python tools/prof_decode.py $mode 65536 $code >> "$OUT" 2>&1
    done
  done
done
cat "$OUT"
