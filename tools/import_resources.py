#!/usr/bin/env python
"""Import the base graphs of the reference's `resources/` directory into this repo's own
edge-list format (JSON), so tests/bench can run where /root/reference is absent (GPU box).

The base graphs are public standard tables (3GPP TS 38.212 Table 5.3.2-3 BG2 set 0;
IEEE 802.16e N=576 R=3/4 "A" code); only the numbers are carried over, as a sparse
per-check edge list  rows[i] = [[col, raw_shift], ...]  (raw shift, used modulo Z).

Run in the build container only:  python tools/import_resources.py [/root/reference/resources]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "neural_ldpc_decoder_torch_b200", "resources")

SOURCES = {
    "nr_bg2_set0": ("basegraph2_set0.txt", "\t", 16),
    "wimax_n576_r34": ("wman_N0576_R34_z24.txt", "\t", 24),
}


def main(src_dir):
    os.makedirs(OUT, exist_ok=True)
    for name, (fname, delim, z) in SOURCES.items():
        bg = np.loadtxt(os.path.join(src_dir, fname), int, delimiter=delim)
        rows = [[[int(j), int(bg[i, j])] for j in range(bg.shape[1]) if bg[i, j] != -1]
                for i in range(bg.shape[0])]
        doc = {"name": name, "M": int(bg.shape[0]), "N": int(bg.shape[1]), "Z_default": z, "rows": rows}
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, bg.shape, sum(len(r) for r in rows), "edges")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/resources")
