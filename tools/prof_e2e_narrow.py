#!/usr/bin/env python
"""chunk-size sweep of the Neural host API with narrow LLR transports (NLDPC_HOST_CHUNK): python tools/prof_e2e_narrow.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200 import neural_ldpc_decoder as nn_  # noqa: E402

B, T = 65536, 10
dev = torch.device("cuda")
bg, Z = load_basegraph("nr_bg2_set0")
m = nn_.NeuralLDPCDecoder(T, B, nn_.ConnectingMatrixTorch(nn_.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)).to(dev)
rs = np.random.RandomState(1)
x = torch.from_numpy((2.0 * (1.2559 * rs.normal(size=(B, bg.shape[1], Z)) - 1.0) / 1.2559 ** 2).astype(np.float32))
inputs = {"fp32": (x.pin_memory(), {}), "fp16": (x.to(torch.float16).pin_memory(), {}),
          "int8": (torch.clamp(torch.round(x / 0.25), -127, 127).to(torch.int8).pin_memory(), {"scale": 0.25})}
for chunk in (2048, 4096, 8192, 16384, 32768):
    os.environ["NLDPC_HOST_CHUNK"] = str(chunk)
    row = []
    for name, (xin, kw) in inputs.items():
        for _ in range(3):
            m.decode_host(xin, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            m.decode_host(xin, **kw)
        dt = (time.perf_counter() - t0) / 10
        row.append(f"{name} {B / dt / 1e6:6.2f} M cw/s ({dt * 1e3:.2f} ms)")
    print(f"chunk {chunk:6d}: " + "   ".join(row))
