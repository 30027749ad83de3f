#!/bin/bash
# A/B of library variants on the headline launch (BG2 z16, B=65536, T=10, packed): tools/exp_ab.sh OUT ROUNDS v1 v2 ...
# ("base" = the in-tree library, anything else = build/exp/<name>.so); variants interleaved within each round.
OUT=$1; ROUNDS=$2; shift 2
: > "$OUT"
for round in $(seq 1 $ROUNDS); do
for v in "$@"; do
  if [ "$v" = base ]; then unset NLDPC_LIB_PATH; else export NLDPC_LIB_PATH=$PWD/build/exp/$v.so; fi
  echo -n "$v " >> "$OUT"; python tools/prof_decode.py packed 65536 nr_bg2_set0 >> "$OUT" 2>&1
done
done
cat "$OUT"
