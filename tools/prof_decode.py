#!/usr/bin/env python
"""Small driver for ncu: a few launches of the decode kernel on the bench workload (BG2 z16, B=65536, T=10).
usage: python tools/prof_decode.py [packed|list|both] [B] [code]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402
import neural_ldpc_decoder_torch_b200.ops  # noqa: E402,F401

mode = sys.argv[1] if len(sys.argv) > 1 else "both"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
code = sys.argv[3] if len(sys.argv) > 3 else "nr_bg2_set0"
T = 10
bg, Z = load_basegraph(code)
N, E = bg.shape[1], int((bg != -1).sum())
dev = torch.device("cuda:0")
cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
gid = cm.graph_id(dev)
rs = np.random.RandomState(0)
w = torch.from_numpy(rs.uniform(0.3, 1.3, (T, E)).astype(np.float32)).to(dev)
b = torch.from_numpy((0.2 * rs.normal(size=(T, E))).astype(np.float32)).to(dev)
sigma = 1.2559 if Z == 16 else 0.62095
g = torch.Generator(device=dev).manual_seed(2042)
xa = (2.0 * (sigma * torch.randn((B, N, Z), generator=g, device=dev) - 1.0) / sigma ** 2).float().contiguous()
for _ in range(4):
    if mode in ("packed", "both"):
        torch.ops.nldpc.neural_hard(xa, w, b, gid, False)
    if mode in ("list", "both"):
        torch.ops.nldpc.neural_forward(xa, w, b, gid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    torch.ops.nldpc.neural_hard(xa, w, b, gid, False) if mode != "list" else torch.ops.nldpc.neural_forward(xa, w, b, gid)
e1.record()
torch.cuda.synchronize()
print(f"{mode} B={B} {code}: {e0.elapsed_time(e1) / 3:.3f} ms/launch -> {B / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e6:.2f} M cw/s")
