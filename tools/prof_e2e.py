#!/usr/bin/env python
"""end-to-end host API timing (pinned host LLRs in, packed decisions out): python tools/prof_e2e.py [B]   (env NLDPC_HOST_CHUNK)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import load_basegraph, ops  # noqa: E402
from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bg, Z = load_basegraph("nr_bg2_set0")
dev = torch.device("cuda:0")
cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
gid = cm.graph_id(dev)
E = int((bg != -1).sum())
xa = torch.randn(B, bg.shape[1], Z).mul_(1.6).sub_(1.27).pin_memory()
w = torch.full((10, E), 0.8)
b = torch.zeros((10, E))
for _ in range(3):
    ops.neural_decode_host(gid, xa, w, b)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter()
    ops.neural_decode_host(gid, xa, w, b)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
print(f"chunk={os.environ.get('NLDPC_HOST_CHUNK', '4096')} B={B}: median {np.median(ts) * 1e3:.3f} ms min {ts.min() * 1e3:.3f} ms -> {B / np.median(ts) / 1e6:.2f} M cw/s")
