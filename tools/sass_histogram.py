#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): the evidence for what the kernels are made of —
bulk TMA (UBLKCP), mbarrier (SYNCS), packed fp32 adds (FADD2), 3-input min (FMNMX3), cp.async (LDGSTS) — and for what they
must NOT contain: FFMA / FFMA2 in the bit-exact decode paths (every multiply and add rounds separately).

    python tools/sass_histogram.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neural_ldpc_decoder_torch_b200", "libnldpc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n   # noqa: E731
kernels, cur, arch = collections.OrderedDict(), None, set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
watch = ["UBLKCP", "SYNCS", "FADD2", "FMNMX3", "FMNMX", "FADD", "FMUL", "FFMA", "FFMA2", "LDGSTS", "LDS", "STS", "LDC", "LDCU", "ATOMS", "RED", "MUFU",
         "F2FP", "HADD2", "BAR", "UTMALDG", "UTCHMMA"]
print(f"library: {os.path.relpath(lib, ROOT)}   cubin archs: {sorted(arch)}   kernels: {len(kernels)}")
print("columns: total instructions, then opcode families (prefix match; '.' = 0)")
print(f"{'total':>7} " + " ".join(f"{w:>7}" for w in watch) + "  kernel")
for name, cnt in kernels.items():
    fam = collections.Counter()
    for op, n in cnt.items():
        base = op.split(".")[0]
        fam[base] += n
    row = [sum(cnt.values())] + [fam.get(w, 0) for w in watch]
    pretty = demangle(name)
    pretty = re.sub(r"nldpc::(gen::)?", "", pretty)
    print(" ".join(f"{v if v else '.':>7}" for v in row) + "  " + pretty[:150])
print()
print("The specialised decode kernels (nldpc_spec_neural_kernel<...>) and the table-driven Neural kernel contain no FFMA / FFMA2: their "
      "arithmetic is compiled with -fmad=false and written with __fadd_rn / __fmul_rn, every multiply and add rounds separately (the "
      "bit-exactness contract).  FFMA appears in the floating-point-tolerance paths only: the SP decoder's tanh / atanh inside "
      "nldpc_generic_boosted_kernel, the loss / optimiser kernels, and the backward sweeps.  UBLKCP = 1-D bulk TMA (cp.async.bulk), "
      "SYNCS = mbarrier, FADD2 = packed add.rn.f32x2, FMNMX3 = 3-input min, LDGSTS = cp.async, MUFU in the training variant = "
      "ex2 / lg2 / rcp of the fused BCE, F2FP = fp16 packing of the training dump.")
