"""helpers for the staging fixtures (tests/golden/staging_*.npz, tools/gen_golden_staging.py): rebuild the module, decode
the recorded forward() calls, and an oracle-side emulation of the reference's stateful forward
(/root/reference/src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py:286-538)."""
import numpy as np
import torch

import oracle


def build_staging_module(d, device="cpu"):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    Z, T = int(d["Z"]), int(d["T"])
    B = (d["c0_xa"] if "c0_xa" in d.files else d["c0_xa_list"][0]).shape[0]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=d["basegraph"]), device=torch.device(device))
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*[int(v) for v in d["sharing"]]),
                                 decoding_type=DecoderType(int(d["decoder_type"])), decoder_qms_qbit=int(d["qbit"]),
                                 fixed_iterative_nodes=[int(v) for v in d["fixed_nodes"]])
    names = {n for n, _ in m.named_parameters()}
    gold = {k[len("param_"):] for k in d.files if k.startswith("param_")}
    assert names == gold, (names ^ gold)           # same parameter set as the reference registers (:139-151)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.copy_(torch.from_numpy(d["param_" + n]))
    return m.to(device)


def call_args(d, k):
    """-> (xa ndarray | list of ndarrays, target_iter, fixed_iter, fixed_iter_weight list | None) of recorded call k"""
    p = f"c{k}_"
    xa = [a for a in d[p + "xa_list"]] if (p + "xa_list") in d.files else d[p + "xa"]
    kind = int(d[p + "target_kind"])
    ti = d[p + "target_iter"]
    target = None if kind == 0 else (int(ti[0]) if kind == 1 else [int(v) for v in ti])
    fixed = [int(v) for v in d[p + "fixed_iter"]] if int(d[p + "has_fixed"]) else None
    fw = None
    if fixed is not None:
        fw = [d[p + f"fw{i}"] for i in range(len(fixed))]
    return xa, target, fixed, fw


class OracleStatefulBoosted:
    """The reference's forward() semantics over oracle.boosted_step: persistent llr / outputs, per-call iteration list,
    fixed_iter weights consumed in sorted-iteration order, list-xa re-assigning the channel input per iteration."""

    def __init__(self, module):
        self.m = module
        g = module.conn_mat.graph
        B, T = module.batch_size, module.iter_node_counts
        self.llr = [np.zeros((B, g.E, module.Z), np.float32) for _ in range(T + 1)]          # [B, E, Z] (oracle layout)
        self.outputs = [np.zeros((B, module.N * module.Z), np.float32) for _ in range(T)]

    def forward(self, xa, target_iter=None, fixed_iter=None, fixed_iter_weight=None):
        m = self.m
        # the reference appends the fixed iterations to the CALLER's target_iter list (:286-296: `iteration = target_iter`,
        # then `iteration.append`), so they are also returned
        iteration = ([target_iter] if isinstance(target_iter, int) else
                     (target_iter if isinstance(target_iter, list) else list(range(m.iter_node_counts))))
        for t in (fixed_iter or []):
            if t not in iteration:
                iteration.append(t)
        iteration = sorted(iteration)
        fixed = list(fixed_iter) if fixed_iter is not None else []
        is_list = isinstance(xa, list)
        if not is_list:
            xin, xo = xa.copy(), xa.copy()
        dec = {"SP": 0, "MS": 1, "QMS": 2}[m.decoding_type.name]
        rng = (float(m.allowed_llr_range.start), float(m.allowed_llr_range.end))
        fidx = 0
        for t in iteration:
            if is_list:
                xin, xo = xa[t].copy(), xa[t].copy()
            fw = None
            if t in fixed:
                fw = [torch.from_numpy(np.asarray(fixed_iter_weight[fidx]))]
            with torch.no_grad():
                vn_w, cn_w, ucn_w, cu, mix = m.fold_weights([t], torch.device("cpu"), fixed, fw)
            npf = lambda w: None if w is None else w[0].detach().numpy()      # noqa: E731
            self.llr[t + 1], self.outputs[t] = oracle.boosted_step(
                m.conn_mat.basegraph, m.Z, dec, int(m.decoder_qms_qbit), rng, xin, xo, self.llr[t], npf(vn_w), npf(cn_w),
                npf(ucn_w), cu, mix, None if t == 0 else self.outputs[t - 1])
            if t in fixed:
                fidx += 1
        if isinstance(target_iter, int):
            return self.outputs[target_iter][None]
        if isinstance(target_iter, list):
            return np.stack([self.outputs[i] for i in target_iter])
        return np.stack(self.outputs)

    def state(self):
        """(outputs [T,B,NZ], llr [T+1,B,Z,E]) in the module's layout"""
        return np.stack(self.outputs), np.stack([l.transpose(0, 2, 1) for l in self.llr])
