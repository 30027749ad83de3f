#!/bin/bash
# Experiment helper: build a variant of libnldpc_b200.so with extra -D flags for some translation units.
#   tools/build_variant.sh NAME "-DFOO=1 -DBAR=2" nldpc_spec.cu nldpc_spec_boosted_bg2.cu
# -> build/exp/NAME.so (all other objects are taken from build/csrc; run `make -C neural_ldpc_decoder_torch_b200/csrc` first).
# Use with NLDPC_LIB_PATH=build/exp/NAME.so.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; DEFS=$2; shift 2
OUT=$ROOT/build/exp/$NAME
mkdir -p "$OUT"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v $NLDPC_EXTRA_NVFLAGS"
OBJS=""
for f in "$ROOT"/build/csrc/*.o; do
  b=$(basename "$f" .o); skip=0
  for tu in "$@"; do [ "$b.cu" = "$tu" ] && skip=1; done
  [ $skip = 0 ] && OBJS="$OBJS $f"
done
pids=""
for tu in "$@"; do
  b=$(basename "$tu" .cu)
  nvcc $FLAGS $DEFS -c -o "$OUT/$b.o" "$ROOT/neural_ldpc_decoder_torch_b200/csrc/$tu" 2> "$OUT/$b.ptxas.log" &
  pids="$pids $!"
  OBJS="$OBJS $OUT/$b.o"
done
for p in $pids; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/build/exp/$NAME.so" $OBJS
echo "built $ROOT/build/exp/$NAME.so"
