#!/usr/bin/env python
"""Fixtures from the LIVE reference for the structure matrices and the checkpoint format (build container only):

  tests/golden/dense_matrix_sha.json   SHA-256 of every dense structure matrix the reference builds for the two built-in
                                       codes (neural ConnectingMatrix.py:68-140, boosted :82-163), as the fp32 C-contiguous
                                       tensors its state_dict() holds, plus shapes and sums
  tests/golden/ref_ckpt_*.pth          checkpoints WRITTEN BY THE REFERENCE's CheckPointUtil.save (CheckPointUtil.py:21-62) for
                                       a toy quasi-cyclic code (M=3, N=6, Z=4: the dense buffers stay a few KB), Neural and
                                       Boosted model with non-trivial weights, optimizer state, epoch, metrics and config

    PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_checkpoint.py
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402
from gen_golden import BoostedNeuralLDPCDecoder, DecoderType, NodeWeightSharingConfig, bref, nref  # noqa: E402
from checkpoint_utils import CheckPointUtil  # noqa: E402  (the reference's)

TOY_BG = np.array([[1, 3, -1, 0, -1, -1],
                   [2, -1, 1, 0, 0, -1],
                   [-1, 0, 3, -1, 2, 0]], dtype=np.int64)
TOY_Z = 4


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32)).tobytes()).hexdigest()


def main():
    doc = {}
    for code, (bg, Z) in gg.GRAPHS.items():
        for fam, mod, T in (("neural", nref, 1), ("boosted", bref, 1)):
            cm = mod.ConnectingMatrixTorch(mod.ConnectingMatrix(Z=Z, basegraph=bg))
            if fam == "neural":
                model = nref.NeuralLDPCDecoder(T, 2, cm)
            else:
                model = BoostedNeuralLDPCDecoder(T, 2, cm)
            entry = {}
            for k, v in model.state_dict().items():
                if k.startswith(("W_", "Lift_")):
                    entry[k] = {"shape": list(v.shape), "sum": float(v.double().sum()), "sha256": sha(v)}
            doc[f"{code}_{fam}"] = entry
            print(code, fam, {k: e["sha256"][:12] for k, e in entry.items()})
    with open(os.path.join(gg.OUT, "dense_matrix_sha.json"), "w") as f:
        json.dump(doc, f, indent=1)

    rs = np.random.RandomState(77)
    with tempfile.TemporaryDirectory() as tmp:
        ck = CheckPointUtil(checkpoint_dir=tmp)
        # Neural, T=3
        cm = nref.ConnectingMatrixTorch(nref.ConnectingMatrix(Z=TOY_Z, basegraph=TOY_BG))
        m = nref.NeuralLDPCDecoder(3, 2, cm)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.from_numpy(rs.uniform(-0.5, 1.5, size=tuple(p.shape)).astype(np.float32)))
        opt = torch.optim.Adam(m.parameters(), lr=1e-3)
        xa = torch.from_numpy(rs.normal(size=(2, 6, TOY_Z)).astype(np.float32))
        loss = sum(o.square().mean() for o in m(xa))
        loss.backward()
        opt.step()
        p1 = ck.save("ref_ckpt_neural_toy.pth", m, optimizer=opt, epoch=7,
                     metrics={"loss": 0.125, "ber_last_iter": 1.5e-3, "fer_last_iter": 2.5e-2}, config={"T": 3, "code": "toy"})
        os.replace(p1, os.path.join(gg.OUT, "ref_ckpt_neural_toy.pth"))
        # Boosted, T=4, cn=2 / ucn=2 / vn=3
        cmb = bref.ConnectingMatrixTorch(bref.ConnectingMatrix(Z=TOY_Z, basegraph=TOY_BG))
        mb = BoostedNeuralLDPCDecoder(4, 2, cmb, node_weight_sharing_config=NodeWeightSharingConfig(2, 2, 3), decoding_type=DecoderType.QMS)
        with torch.no_grad():
            for p in mb.parameters():
                p.copy_(torch.from_numpy(rs.uniform(0.2, 1.8, size=tuple(p.shape)).astype(np.float32)))
        optb = torch.optim.Adam(mb.get_trainable_parameters(), lr=1e-3)
        p2 = ck.save("ref_ckpt_boosted_toy.pth", mb, optimizer=optb, epoch=2, metrics={"loss": 0.5}, config={"T": 4})
        os.replace(p2, os.path.join(gg.OUT, "ref_ckpt_boosted_toy.pth"))
    for f in ("ref_ckpt_neural_toy.pth", "ref_ckpt_boosted_toy.pth"):
        print(f, os.path.getsize(os.path.join(gg.OUT, f)), "bytes")
    np.savez(os.path.join(gg.OUT, "ref_ckpt_toy_graph.npz"), basegraph=TOY_BG, Z=TOY_Z)


if __name__ == "__main__":
    main()
