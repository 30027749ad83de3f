#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the LIVE, UNMODIFIED reference on CPU.

Only runs in the build container (needs /root/reference).  The fixtures pin the oracle
(oracle/nldpc_oracle.c) and, through it, the CUDA path.  Re-run:

    PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden.py

Every case stores its inputs, parameters and the reference outputs (fp32, bit patterns
preserved by .npz); "hash-only" cases store a SHA-256 of the output bytes for larger batches
whose inputs are reproducible from a numpy seed.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("NLDPC_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "src"))

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
RES = os.path.join(REF, "resources")

import neural_ldpc_decoder as nref  # noqa: E402
import boosted_neural_ldpc_decoder as bref  # noqa: E402
from boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss  # noqa: E402
from boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from boosted_neural_ldpc_decoder.struct.LossType import LossType  # noqa: E402
from boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402

GRAPHS = {
    "bg2": (np.loadtxt(os.path.join(RES, "basegraph2_set0.txt"), int, delimiter="\t"), 16),
    "wimax": (np.loadtxt(os.path.join(RES, "wman_N0576_R34_z24.txt"), int, delimiter="\t"), 24),
}
GEN_BG2 = np.loadtxt(os.path.join(RES, "gen_matrix_bg2_z16.txt"), int, delimiter=",")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def neural_model(code, T, B):
    bg, Z = GRAPHS[code]
    cm = nref.ConnectingMatrixTorch(nref.ConnectingMatrix(Z=Z, basegraph=bg))
    return nref.NeuralLDPCDecoder(T, B, cm), bg, Z


def neural_inputs(code, B, snr=2.0):
    """SURVEY Appendix D1 recipe: reference neural datagen, seeds 2042/1074, all-zero codeword."""
    bg, Z = GRAPHS[code]
    M, N = bg.shape
    gen = GEN_BG2 if code == "bg2" else np.zeros(((N - M) * Z, N * Z), dtype=np.int64)
    dg = nref.AWGNPassedDatagen(N=N, M=M, snr_db=np.array([snr]), awgn_noise_seed=2042, wordgen_random_seed=1074,
                                gen_matrix=gen)
    x, _ = dg(word_length=B, Z=Z, is_y_all_zero=True)
    return np.reshape(x[0], [B, N, Z]).astype(np.float32)


def run_neural(code, T, xa, w=None, b=None):
    model, bg, Z = neural_model(code, T, xa.shape[0])
    with torch.no_grad():
        if w is not None:
            for t in range(T):
                model.weights_var[t].copy_(torch.from_numpy(w[t]))
                model.biases_var[t].copy_(torch.from_numpy(b[t]))
        outs = model(torch.from_numpy(xa))
    return np.stack([o.numpy() for o in outs]).astype(np.float32)


def trained_like(T, E, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = (0.3 + torch.rand(T, E, generator=g)).numpy().astype(np.float32)
    b = (0.2 * torch.randn(T, E, generator=g)).numpy().astype(np.float32)
    return w, b


def gen_neural():
    for code in ("bg2", "wimax"):
        bg, Z = GRAPHS[code]
        E = int((bg != -1).sum())
        T = 10
        # D1: init weights, B=8
        xa = neural_inputs(code, 8)
        w = np.full((T, E), 0.5, np.float32)
        b = np.zeros((T, E), np.float32)
        out = run_neural(code, T, xa)
        np.savez_compressed(os.path.join(OUT, f"neural_{code}_init.npz"), xa=xa, w=w, b=b, out=out, Z=Z, basegraph=bg)
        print(code, "init", sha(xa)[:16], sha(out[0])[:16], sha(out[9])[:16], int((out[9] < 0).sum()))
        # trained-like weights, B=4, plus adversarial inputs: exact zeros (punctured columns), ties, huge values
        xa = neural_inputs(code, 4).copy()
        xa[1, :2, :] = 0.0            # punctured columns -> exact-zero v2c at iteration 0 (10000 mask path)
        xa[2, 3, :] = 0.0
        xa[2, 5, ::2] = -0.0
        xa[3] = np.round(xa[3] * 2) / 2   # grid values -> ties in the min
        xa[3, 0, 0] = 30000.0             # above the 10000 cap
        xa[3, 1, :] = 20000.0
        w, b = trained_like(T, E, seed=0)
        out = run_neural(code, T, xa, w, b)
        np.savez_compressed(os.path.join(OUT, f"neural_{code}_trained.npz"), xa=xa, w=w, b=b, out=out, Z=Z, basegraph=bg)
        print(code, "trained", sha(out[9])[:16])
    # hash-only larger batches (inputs reproducible from numpy seed): BG2 B=64 and WiMAX B=128, T=10
    hashes = {}
    for code, B in (("bg2", 64), ("wimax", 128)):
        bg, Z = GRAPHS[code]
        M, N = bg.shape
        E = int((bg != -1).sum())
        rs = np.random.RandomState(1234)
        sigma = 1.2559 if code == "bg2" else 0.62095
        xa = (2.0 * (sigma * rs.normal(0, 1, (B, N, Z)) - 1.0) / sigma ** 2).astype(np.float32)
        w, b = trained_like(10, E, seed=1)
        out = run_neural(code, 10, xa, w, b)
        hashes[code] = {"B": B, "T": 10, "seed": 1234, "sigma": sigma, "wb_seed": 1,
                        "w_sha": sha(w), "b_sha": sha(b), "xa_sha": sha(xa),
                        "out_sha": [sha(out[t]) for t in range(10)],
                        "packed_sha": sha(np.packbits(out[9] < 0, axis=1, bitorder="little"))}
        # weights are torch-RNG generated: store them (small)
        np.savez_compressed(os.path.join(OUT, f"neural_{code}_hash_wb.npz"), w=w, b=b)
        print(code, "hash", hashes[code]["out_sha"][9][:16])
    with open(os.path.join(OUT, "neural_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)


def boosted_inputs(code, B, dtype, qbit, y_all_zero):
    """SURVEY Appendix D2 recipe: reference boosted datagen, mix_snr, seeds 2042/1074."""
    bg, Z = GRAPHS[code]
    M, N = bg.shape
    dg = bref.AWGNPassedDatagen(N=N, M=M, snr_db=np.array([2, 2.5, 3.0, 3.5, 4.0]), awgn_noise_seed=2042,
                                wordgen_random_seed=1074, gen_matrix=GEN_BG2 if code == "bg2" else None)
    x, y = dg(gentype="mix_snr", word_length=B, Z=Z, is_y_all_zero=y_all_zero, decoding_type=dtype,
              decoder_qms_qbit=qbit)
    return np.reshape(x, [B, N, Z]).astype(np.float32), np.asarray(y).astype(np.float32)


def boosted_model(code, T, B, sharing, dtype, qbit, rng, fill=None, fixed_nodes=()):
    bg, Z = GRAPHS[code]
    cm = bref.ConnectingMatrixTorch(bref.ConnectingMatrix(Z=Z, basegraph=bg))
    model = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing),
                                     decoding_type=dtype, decoder_qms_qbit=qbit,
                                     fixed_iterative_nodes=list(fixed_nodes))
    params = {}
    with torch.no_grad():
        for name, p in model.named_parameters():
            if fill is not None:
                p.fill_(fill)
            else:
                p.copy_(torch.from_numpy(rng.uniform(0.4, 1.3, size=tuple(p.shape)).astype(np.float32)))
            params[name] = p.detach().numpy().copy()
    return model, params


def gen_boosted():
    cases = []
    # (code, sharing, decoder, qbit, T, B, fill, all_zero)
    for code in ("bg2", "wimax"):
        for sharing in ((3, 0, 0), (3, 0, 3)):
            cases.append((code, sharing, DecoderType.QMS, 5, 20, 8, 0.75, code != "bg2"))   # Appendix D2
    rnd = [
        ("wimax", (0, 0, 0), DecoderType.QMS, 5, 6), ("wimax", (1, 0, 2), DecoderType.QMS, 5, 6),
        ("wimax", (2, 0, 3), DecoderType.QMS, 5, 6), ("wimax", (1, 1, 0), DecoderType.QMS, 5, 6),
        ("wimax", (2, 2, 2), DecoderType.QMS, 5, 6), ("wimax", (3, 3, 3), DecoderType.QMS, 5, 6),
        ("wimax", (3, 1, 0), DecoderType.QMS, 5, 6),   # unequal UCN: indicator computed, CN-only weights
        ("wimax", (3, 0, 3), DecoderType.MS, 5, 6), ("wimax", (1, 1, 2), DecoderType.MS, 5, 6),
        ("bg2", (3, 0, 3), DecoderType.MS, 5, 6), ("bg2", (3, 3, 3), DecoderType.QMS, 5, 6),
        ("bg2", (1, 0, 0), DecoderType.QMS, 6, 5), ("bg2", (2, 0, 2), DecoderType.QMS, -5, 5),
        ("wimax", (3, 0, 3), DecoderType.QMS, 4, 5), ("wimax", (3, 0, 3), DecoderType.QMS, 3, 5),
        ("wimax", (3, 0, 3), DecoderType.QMS, 7, 5),   # unknown q_bit -> no quantisation, no clamp on messages
        ("wimax", (3, 0, 3), DecoderType.SP, 5, 5), ("bg2", (1, 0, 2), DecoderType.SP, 5, 5),
    ]
    for code, sharing, dt, q, T in rnd:
        cases.append((code, sharing, dt, q, T, 4, None, False if code == "bg2" else True))
    index = []
    for idx, (code, sharing, dt, q, T, B, fill, all_zero) in enumerate(cases):
        bg, Z = GRAPHS[code]
        rng = np.random.RandomState(100 + idx)
        xa, y = boosted_inputs(code, B, dt, q, all_zero)
        model, params = boosted_model(code, T, B, sharing, dt, q, rng, fill)
        with torch.no_grad():
            outs = model(torch.from_numpy(xa))
        out = np.stack([o.numpy() for o in outs]).astype(np.float32)
        llr_last = model.llr[T].numpy().astype(np.float32)   # [B, Z, E]
        name = f"boosted_{idx:02d}_{code}_{dt.name}_q{q}_cn{sharing[0]}ucn{sharing[1]}vn{sharing[2]}"
        np.savez_compressed(os.path.join(OUT, name + ".npz"), xa=xa, y=y, out=out, llr_last=llr_last, Z=Z, basegraph=bg,
                            sharing=np.array(sharing), decoder_type=dt.value, qbit=q, T=T,
                            **{"param_" + k: v for k, v in params.items()})
        index.append(name)
        print(name, sha(xa)[:16], sha(out[0])[:16], sha(out[-1])[:16], int((out[-1] > 0).sum()))
    with open(os.path.join(OUT, "boosted_index.json"), "w") as f:
        json.dump(index, f, indent=1)


def gen_train():
    """One training step as in train/train_BoostedNeuralLDPCDecoder.py:270-294 (loss + grads), Appendix D4,
    plus a Neural grad case (autograd of the reference) for the backward kernel."""
    # Boosted: BG2, B=20, T=20, QMS5, cn3/vn3, init weights 1, random codewords, BCE etha=1
    for tag, sharing, dt, etha, fill in (("d4", (3, 0, 3), DecoderType.QMS, 1.0, None),
                                         ("cn1vn2_ms", (1, 0, 2), DecoderType.MS, 1.3, "rand"),
                                         ("cn2vn3_qms", (2, 0, 3), DecoderType.QMS, 1.3, "rand")):
        code, T, B = "bg2", 20 if tag == "d4" else 6, 20 if tag == "d4" else 6
        bg, Z = GRAPHS[code]
        xa, y = boosted_inputs(code, B, dt, 5, False)
        rng = np.random.RandomState(7)
        model, params = boosted_model(code, T, B, sharing, dt, 5, rng, 1.0 if fill is None else None)
        crit = LDPCDecoderLoss(loss_type=LossType.BCE, etha=etha)
        outs = model(torch.from_numpy(xa), target_iter=list(range(T)))
        loss = crit(outs, torch.from_numpy(y), coeff_param=list(range(T)))
        loss.backward()
        grads = {k: p.grad.numpy().copy() for k, p in model.named_parameters()}
        np.savez_compressed(os.path.join(OUT, f"train_boosted_{tag}.npz"), xa=xa, y=y, Z=Z, basegraph=bg,
                            sharing=np.array(sharing), decoder_type=dt.value, qbit=5, T=T, etha=etha,
                            loss=np.float32(loss.item()), loss64=np.float64(loss.item()),
                            **{"param_" + k: v for k, v in params.items()},
                            **{"grad_" + k: v for k, v in grads.items()})
        gn = float(np.sqrt(sum((g.astype(np.float64) ** 2).sum() for g in grads.values())))
        print("train", tag, loss.item(), gn, {k: float(v.reshape(-1)[0]) for k, v in list(grads.items())[:2]})
    # Neural: BG2 B=4 T=5 and WiMAX, BCE on every iteration output (sum), random weights
    for code in ("bg2", "wimax"):
        bg, Z = GRAPHS[code]
        E = int((bg != -1).sum())
        T, B = 5, 4
        xa = neural_inputs(code, B, snr=3.0)
        w, b = trained_like(T, E, seed=3)
        model, _, _ = neural_model(code, T, B)
        with torch.no_grad():
            for t in range(T):
                model.weights_var[t].copy_(torch.from_numpy(w[t]))
                model.biases_var[t].copy_(torch.from_numpy(b[t]))
        outs = model(torch.from_numpy(xa))
        y = torch.zeros(B, outs[0].shape[1])
        crit = LDPCDecoderLoss(loss_type=LossType.BCE, etha=1.2)
        loss = crit(outs, y, coeff_param=list(range(T)))
        loss.backward()
        gw = np.stack([p.grad.numpy() for p in model.weights_var])
        gb = np.stack([p.grad.numpy() for p in model.biases_var])
        np.savez_compressed(os.path.join(OUT, f"train_neural_{code}.npz"), xa=xa, w=w, b=b, Z=Z, basegraph=bg,
                            etha=1.2, loss=np.float32(loss.item()), grad_w=gw, grad_b=gb)
        print("train neural", code, loss.item(), float(np.abs(gw).max()), float(np.abs(gb).max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["neural", "boosted", "train"]
    if "neural" in which:
        gen_neural()
    if "boosted" in which:
        gen_boosted()
    if "train" in which:
        gen_train()
