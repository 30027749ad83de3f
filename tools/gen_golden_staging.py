#!/usr/bin/env python
"""Generate tests/golden/staging_*.npz by running the LIVE, UNMODIFIED reference on CPU: the Boosted staging features of
/root/reference/src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py — sharing type 4 with `fixed_iterative_nodes`
(:139-145, :225-235), `fixed_iter` + `fixed_iter_weight` (:293-312, :327-334, :498-503, :528-531), list-`xa` (:300-303,
:321-323) and staged `target_iter` call SEQUENCES on one module (the stateful `self.llr` / `self.outputs`, :94-101, :512).

Only runs in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tools/gen_golden_staging.py

Each fixture holds the constructor arguments, the parameters, and a list of forward() calls with, per call, its arguments,
the tensors it returned and the module's public state (`outputs`, `llr`) after it.
"""
import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  (sets up the reference import path)
from gen_golden import DecoderType, boosted_inputs, boosted_model, sha  # noqa: E402

OUT = gg.OUT


def record_call(model, store, k, xa, target_iter, fixed_iter, fixed_iter_weight):
    """run one forward() of the reference and store arguments, result and state under prefix c{k}_"""
    p = f"c{k}_"
    if isinstance(xa, list):
        store[p + "xa_list"] = np.stack(xa)
        xin = [torch.from_numpy(a) for a in xa]
    else:
        store[p + "xa"] = xa
        xin = torch.from_numpy(xa)
    store[p + "target_iter"] = np.array([-1] if target_iter is None else
                                        ([target_iter] if isinstance(target_iter, int) else target_iter), dtype=np.int64)
    store[p + "target_kind"] = np.int64(0 if target_iter is None else (1 if isinstance(target_iter, int) else 2))
    store[p + "fixed_iter"] = np.array([] if fixed_iter is None else fixed_iter, dtype=np.int64)
    store[p + "has_fixed"] = np.int64(fixed_iter is not None)
    if fixed_iter_weight is not None:
        for i, w in enumerate(fixed_iter_weight):
            store[p + f"fw{i}"] = w
    with torch.no_grad():
        ret = model(xin, target_iter=None if target_iter is None else (target_iter if isinstance(target_iter, int) else list(target_iter)),
                    fixed_iter=None if fixed_iter is None else list(fixed_iter),
                    fixed_iter_weight=None if fixed_iter_weight is None else [torch.from_numpy(w) for w in fixed_iter_weight])
    if isinstance(ret, torch.Tensor):
        store[p + "ret"] = ret.numpy().astype(np.float32)[None]
    else:
        store[p + "ret"] = np.stack([o.numpy() for o in ret]).astype(np.float32)
    store[p + "outputs"] = np.stack([o.numpy() for o in model.outputs]).astype(np.float32)        # [T, B, NZ]
    store[p + "llr"] = np.stack([l.numpy() for l in model.llr]).astype(np.float32)                # [T+1, B, Z, E]


def main():
    index = []
    rs = np.random.RandomState(4242)

    def case(name, code, sharing, dt, q, T, B, fixed_nodes, calls, all_zero=None):
        bg, Z = gg.GRAPHS[code]
        E = int((bg != -1).sum())
        all_zero = (code != "bg2") if all_zero is None else all_zero
        xa, _ = boosted_inputs(code, B, dt, q, all_zero)
        # a second, different input batch for calls that change the data between stages
        xb = np.roll(xa, 1, axis=0).copy()
        xb = (-xb[:, ::-1, :]).copy() if code == "wimax" else np.roll(xb, 3, axis=2).copy()
        model, params = boosted_model(code, T, B, sharing, dt, q, np.random.RandomState(500 + len(index)), None, fixed_nodes)
        store = dict(Z=Z, basegraph=bg, sharing=np.array(sharing), decoder_type=dt.value, qbit=q, T=T,
                     fixed_nodes=np.array(list(fixed_nodes), dtype=np.int64), n_calls=len(calls),
                     **{"param_" + k: v for k, v in params.items()})
        for k, c in enumerate(calls):
            src = c.get("xa", "a")
            if src == "list":
                n = c["n_list"]
                x = [(np.roll(xa, i, axis=0) if i % 2 == 0 else np.roll(xb, i, axis=0)).copy() for i in range(n)]
            else:
                x = xa if src == "a" else xb
            fw = None
            if c.get("fixed_iter") is not None:
                fw = [rs.uniform(0.4, 1.3, size=c.get("fw_shape", (E,))).astype(np.float32) for _ in c["fixed_iter"]]
            record_call(model, store, k, x, c.get("target_iter"), c.get("fixed_iter"), fw)
        fname = f"staging_{len(index):02d}_{name}"
        np.savez_compressed(os.path.join(OUT, fname + ".npz"), **store)
        index.append(fname)
        last = store[f"c{len(calls) - 1}_outputs"]
        print(fname, sha(last)[:16], [tuple(store[f"c{k}_ret"].shape) for k in range(len(calls))])

    QMS, MS = DecoderType.QMS, DecoderType.MS
    # sharing type 4: per-edge weights at iteration 0 and at the fixed iterative nodes; fetch_param picks the latest node <= t
    case("cn4_fixednodes", "wimax", (4, 0, 0), QMS, 5, 6, 4, [2, 4], [dict()])
    # ... the same with external weights at iterations 1 and 3 (fixed_iter / fixed_iter_weight)
    case("cn4_fixediter", "wimax", (4, 0, 0), QMS, 5, 6, 4, [2, 4], [dict(fixed_iter=[1, 3])])
    case("cn4vn3_ms_fixediter", "bg2", (4, 0, 3), MS, 5, 5, 3, [1], [dict(target_iter=[0, 1, 2, 3, 4], fixed_iter=[0, 2])])
    case("cn4_nonodes_q6", "bg2", (4, 0, 0), QMS, 6, 4, 3, [], [dict()])
    # list-xa: one input tensor per iteration (fixed_iter must be a list, :302)
    case("listxa_cn3vn3", "wimax", (3, 0, 3), QMS, 5, 5, 4, [], [dict(xa="list", n_list=5, fixed_iter=[])])
    case("listxa_cn1ucn1vn2", "wimax", (1, 1, 2), QMS, 5, 4, 4, [], [dict(xa="list", n_list=4, fixed_iter=[])])
    # staged target_iter sequences on ONE module: every call continues from self.llr / self.outputs of the previous ones
    case("staged_bg2_qms333", "bg2", (3, 3, 3), QMS, 5, 8, 4, [],
         [dict(target_iter=[0, 1, 2]), dict(target_iter=[3, 4, 5, 6]), dict(target_iter=7)])
    # ... with the data changing between stages and a late single iteration reading the state another batch left behind
    case("staged_wimax_ms202_stale", "wimax", (2, 0, 2), MS, 5, 6, 4, [],
         [dict(target_iter=[0, 1]), dict(target_iter=[2, 3]), dict(xa="b"), dict(target_iter=4), dict(target_iter=[1, 5])])
    # fixed_iter names an iteration outside target_iter: it is executed too (:293-296), from never-written (zero) state
    case("fixediter_outside_target", "wimax", (3, 0, 0), QMS, 5, 6, 4, [], [dict(target_iter=[0, 1, 2], fixed_iter=[4], fw_shape=(1,))])
    # non-consecutive target list on a fresh module: iteration 2 reads the zero-initialised self.llr[2]; VN scaling compounds
    case("gap_cn3vn3", "wimax", (3, 0, 3), QMS, 5, 5, 4, [], [dict(target_iter=[0, 2, 3])])
    case("gap_ucn_ms", "bg2", (2, 2, 2), MS, 5, 5, 3, [], [dict(target_iter=[1, 3]), dict(target_iter=[0, 1, 2, 3, 4]), dict(xa="b", target_iter=[2, 4])])
    with open(os.path.join(OUT, "staging_index.json"), "w") as f:
        json.dump(index, f, indent=1)


def gen_staged_training():
    """staged training as train/train_BoostedNeuralLDPCDecoder.py:139-181, 270-294 runs it with fixed_iter > 0: iterations
    [0, t0) once under no_grad (the validation pass leaves such graph-less state in self.llr / self.outputs), then
    forward(target_iter=range(t0, T)) -> LDPCDecoderLoss -> backward: the stored state is a constant of the call."""
    from gen_golden import LDPCDecoderLoss, LossType
    for tag, code, sharing, dt, T, B, t0 in (("bg2_qms303", "bg2", (3, 0, 3), DecoderType.QMS, 6, 6, 3),
                                             ("wimax_ms112", "wimax", (1, 1, 2), DecoderType.MS, 5, 5, 2),
                                             ("bg2_qms100", "bg2", (1, 0, 0), DecoderType.QMS, 5, 4, 1)):
        bg, Z = gg.GRAPHS[code]
        xa, y = boosted_inputs(code, B, dt, 5, False if code == "bg2" else True)
        model, params = boosted_model(code, T, B, sharing, dt, 5, np.random.RandomState(900 + T), None, [])
        with torch.no_grad():
            model(torch.from_numpy(xa), target_iter=list(range(t0)))
        outs = model(torch.from_numpy(xa), target_iter=list(range(t0, T)))
        loss = LDPCDecoderLoss(loss_type=LossType.BCE, etha=1.2)(outs, torch.from_numpy(y), coeff_param=list(range(len(outs))))
        loss.backward()
        grads = {k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
        np.savez_compressed(os.path.join(OUT, f"train_staged_{tag}.npz"), xa=xa, y=y, Z=Z, basegraph=bg, sharing=np.array(sharing),
                            decoder_type=dt.value, qbit=5, T=T, t0=t0, etha=1.2, loss=np.float32(loss.item()),
                            out=np.stack([o.detach().numpy() for o in outs]).astype(np.float32),
                            **{"param_" + k: v for k, v in params.items()}, **{"grad_" + k: v for k, v in grads.items()})
        print("train_staged", tag, loss.item(), sorted(grads.keys()))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["calls", "train"]
    if "calls" in which:
        main()
    if "train" in which:
        gen_staged_training()
