#!/bin/bash
# Offline "install" of the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot) so
# that bench.py can time the stock PyTorch CPU path beside the C port (cpu_baseline_stock).
#
# The contract's recipe — pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
# /root/reference — fails in this image: the reference's build backend is hatchling, which is neither installed nor in the
# wheelhouse.  The reference is pure Python and its wheel target lists exactly three packages (pyproject.toml,
# [tool.hatch.build.targets.wheel].packages), so the wheel pip would have unpacked is those three directories: this script
# tries pip first and, when that fails, places them itself.  Nothing under baseline/_ref is product source or ever committed.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${NLDPC_REFERENCE:-/root/reference}
DST=$ROOT/baseline/_ref
[ -d "$REF/src" ] || { echo "no reference at $REF"; exit 1; }
rm -rf "$DST"; mkdir -p "$DST"
TMP=$(mktemp -d); cp -r "$REF" "$TMP/ref"
if python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP/ref" 2> "$TMP/pip.log"; then
  echo "pip install ok"
else
  echo "pip install failed ($(grep -m1 -o "No module named '[a-z]*'" "$TMP/pip.log" || echo see log)); unpacking the wheel's package list by hand"
  for p in neural_ldpc_decoder boosted_neural_ldpc_decoder checkpoint_utils; do cp -r "$REF/src/$p" "$DST/$p"; done
fi
find "$DST" -name __pycache__ -prune -exec rm -rf {} +
rm -rf "$TMP"
ls "$DST"
