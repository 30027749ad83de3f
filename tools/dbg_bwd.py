"""Debug aid: specialised vs table-driven backward sweep on the golden training fixtures (GPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import load_golden  # noqa: E402
from test_neural_gpu import make_model  # noqa: E402


def run(code, generic):
    os.environ["NLDPC_FORCE_GENERIC"] = "1" if generic else "0"
    d = load_golden(f"train_neural_{code}")
    T, B = d["w"].shape[0], d["xa"].shape[0]
    m = make_model(d["basegraph"], int(d["Z"]), T, B, d["w"], d["b"])
    outs = m(torch.from_numpy(d["xa"]).cuda())
    y = torch.zeros(B, outs[0].shape[1], device="cuda")
    loss = sum(torch.nn.functional.binary_cross_entropy_with_logits(-o, y) for o in outs)
    loss.backward()
    torch.cuda.synchronize()
    gw = np.stack([p.grad.cpu().numpy() for p in m.weights_var])
    gb = np.stack([p.grad.cpu().numpy() for p in m.biases_var])
    return gw, gb


for code in ("bg2", "wimax"):
    gs, bs = run(code, False)
    gg, bg = run(code, True)
    print(code, "B,T", load_golden(f"train_neural_{code}")["xa"].shape[0], gs.shape, "spec |gw|", np.abs(gs).max(), "generic |gw|", np.abs(gg).max(),
          "max diff", np.abs(gs - gg).max(), np.abs(bs - bg).max())
