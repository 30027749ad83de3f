"""BoostedNeuralLDPCDecoder — drop-in for
/root/reference/src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py:14-538 (Kwak et al. boosted neural min-sum).

Same constructor, parameter names / shapes / init values (`weight_CN_{t}`, `weight_UCN_{t}`, `weight_VN_{t}`; no bias
parameters are ever created, as in the reference), `fetch_param`, `get_trainable_parameters`, `_apply_constraints`,
forward signature and return convention (the module's own `self.outputs` list / one tensor / a sub-list), and the
`state_dict()` layout incl. the dense structure buffers.

What changes: the loop body (:320-531) runs as ONE CUDA launch per maximal run of consecutive iterations
(`torch.ops.nldpc.boosted_forward`); the weight-sharing types are folded into per-iteration rows by ordinary
autograd-visible indexing/expansion, so gradients reach the actual parameters.

Statefulness (:94-101): `self.outputs[t]` and `self.llr[t+1]` are written for every executed
iteration, as in the reference (:512, :523) — the kernel exports the c2v messages of all iterations of a run from the same
launch (`store_llr = "all"`, default), so partial `target_iter` calls read exactly the state the reference would read
(zeros where nothing was ever stored).  `store_llr = "last"` keeps only the state after the last iteration of a run and
`"none"` skips the export (12.6 KB per BG2 codeword-iteration): throughput settings; with them a call that would continue
from an iteration whose state was not stored raises instead of reading stale data.
Under `torch.no_grad()` the state tensors are `[B, Z, :E]` views of rows padded to a multiple of 4 floats (BG2: 197 of 200;
WiMAX's 88 needs none) — same shape and values as the reference's, not contiguous when padded (`.reshape` / `.contiguous()`
where a flat view is wanted): 16-byte rows are what lets the kernels write them with vector stores.
"""
from typing import Optional

import torch
import torch.nn as nn

from .. import ops
from .ConnectingMatrixTorch import ConnectingMatrixTorch
from .struct.Clipping import Clipping
from .struct.DecoderType import DecoderType
from .struct.NodeType import NodeType
from .struct.NodeWeightSharingConfig import NodeWeightSharingConfig
from .struct.ParamType import ParamType

_BUFFERS = (("W_odd2even", "W_odd2even"), ("W_skipconn2even", "W_skipconn2even"), ("W_even2odd", "W_even2odd"),
            ("W_even2odd_with_self", "W_even2odd_with_self"), ("W_output", "W_output"), ("W_skipconn2odd", "W_skipconn2odd"),
            ("Lift_Matrix1", "lifting_matrix_1"), ("Lift_Matrix2", "lifting_matrix_2"))


class BoostedNeuralLDPCDecoder(nn.Module):
    def __init__(
            self,
            iter_node_counts,
            batch_size,
            connecting_matrix: ConnectingMatrixTorch,
            node_weight_sharing_config: NodeWeightSharingConfig = NodeWeightSharingConfig(
                cn_weight_sharing=3,
                ucn_weight_sharing=0,
                vn_weight_sharing=0,
            ),
            decoding_type: DecoderType = DecoderType.QMS,
            decoder_qms_qbit: int = 5,
            fixed_iterative_nodes: list[int] = [],
            fixed_iterative_nodes_init_weight: int = 0,
            allowed_weight_range: Clipping = Clipping(start=0, end=2),
            allowed_bias_range: Clipping = Clipping(start=0, end=2),
            allowed_llr_range: Clipping = Clipping(abs=20.0),
            dtype_cn_weight: torch.dtype = torch.float32,
            dtype_ucn_weight: torch.dtype = torch.float32,
            dtype_vn_weight: torch.dtype = torch.float32,
            init_cn_weight: float = 1,
            init_ucn_weight: float = 1,
            init_vn_weight: float = 1,
            dtype_cn_bias: torch.dtype = torch.float32,
            dtype_ucn_bias: torch.dtype = torch.float32,
            dtype_vn_bias: torch.dtype = torch.float32,
            init_cn_bias: float = 1,
            init_ucn_bias: float = 1,
            init_vn_bias: float = 1,
    ):
        super(BoostedNeuralLDPCDecoder, self).__init__()
        for dt in (dtype_cn_weight, dtype_ucn_weight, dtype_vn_weight):
            if dt != torch.float32:
                raise ValueError("the B200 decode path computes in float32; weight dtypes other than float32 are not supported")
        self.iter_node_counts = iter_node_counts
        self.batch_size = batch_size
        self.conn_mat = connecting_matrix
        self.N, self.M, self.Z = self.conn_mat.N, self.conn_mat.M, self.conn_mat.Z
        self.sum_edge = self.conn_mat.sum_edge
        self.neurons_per_odd_layer = self.conn_mat.neurons_per_odd_layer
        self.neurons_per_even_layer = self.conn_mat.neurons_per_even_layer

        self.node_weight_sharing_config = node_weight_sharing_config
        self.decoding_type = decoding_type
        self.decoder_qms_qbit = decoder_qms_qbit
        self.fixed_iterative_nodes = fixed_iterative_nodes
        self.fixed_iterative_nodes_init_weight = fixed_iterative_nodes_init_weight
        self.allowed_weight_range = allowed_weight_range
        self.allowed_bias_range = allowed_bias_range
        self.allowed_llr_range = allowed_llr_range
        self.dtype_cn_weight, self.dtype_ucn_weight, self.dtype_vn_weight = dtype_cn_weight, dtype_ucn_weight, dtype_vn_weight
        self.init_cn_weight, self.init_ucn_weight, self.init_vn_weight = init_cn_weight, init_ucn_weight, init_vn_weight
        self.dtype_cn_bias, self.dtype_ucn_bias, self.dtype_vn_bias = dtype_cn_bias, dtype_ucn_bias, dtype_vn_bias
        self.init_cn_bias, self.init_ucn_bias, self.init_vn_bias = init_cn_bias, init_ucn_bias, init_vn_bias

        # public state of the reference (:94-101)
        dev = self.conn_mat.device
        E = int(self.sum_edge)
        self.outputs = [torch.zeros((self.batch_size, self.N * self.Z), dtype=torch.float32, device=dev)
                        for _ in range(self.iter_node_counts)]
        self.llr = [torch.zeros((self.batch_size, self.Z, E), dtype=torch.float32, device=dev)
                    for _ in range(self.iter_node_counts + 1)]
        # the zero-initialised state IS the reference's state; entries only become invalid when a run executed an iteration
        # without storing its messages (store_llr "last" / "none")
        self._llr_valid = [True] * (self.iter_node_counts + 1)
        self.store_llr = "all"                                         # "all" (reference state) | "last" | "none"

        self._erow_cache = {}
        self._register_params()
        self._flatten_params()
        self._register_state_dict_hook(_add_dense_buffers)
        self.register_load_state_dict_pre_hook(_drop_dense_buffers)

    # ---- parameters (same names and shapes as the reference, :105-151) -------------------------------------
    def _param_name(self, param_type: ParamType, node_type: NodeType, iterative_node_identifier: int):
        return f"{param_type.value}_{node_type.value}_{iterative_node_identifier}"

    def _iterations_with_params(self, sharing_type):
        if sharing_type in [1, 2, 3]:
            return list(range(self.iter_node_counts))
        its = [0]
        if self.fixed_iterative_nodes is not None:
            its += list(self.fixed_iterative_nodes)
        return its

    def _register_params(self):
        init = {NodeType.CN: self.init_cn_weight, NodeType.UCN: self.init_ucn_weight, NodeType.VN: self.init_vn_weight}
        for node_type, sharing_type in self.node_weight_sharing_config:
            if sharing_type == 0:
                continue
            if sharing_type in [1, 4]:
                shape = (int(self.sum_edge),)
            elif sharing_type in [2, 5]:
                shape = (self.M,) if node_type in [NodeType.CN, NodeType.UCN] else (self.N,)
            elif sharing_type == 3:
                shape = (1,)
            else:
                raise ValueError(f"Unsupported sharing type {sharing_type} for {node_type}")
            for iteration in self._iterations_with_params(sharing_type):
                name = self._param_name(ParamType.Weight, node_type, iteration)
                setattr(self, name, nn.Parameter(torch.full(shape, init[node_type], dtype=torch.float32)))

    def _flatten_params(self):
        from .._flatparams import flatten_
        flatten_(list(self.parameters()))

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)          # .to() / .cuda() give every parameter its own storage again
        self._flatten_params()
        return out

    def _apply_constraints(self):
        """clamp weights into allowed_weight_range after an optimiser step (:153-179)"""
        for node_type, sharing_type in self.node_weight_sharing_config:
            if sharing_type == 0:
                continue
            if sharing_type in [1, 2, 3]:
                iterations = range(self.iter_node_counts)
            elif self.fixed_iterative_nodes is not None and len(self.fixed_iterative_nodes) > 0:
                iterations = self.fixed_iterative_nodes
            else:
                iterations = [0]
            for iteration in iterations:
                for param_type, rng in ((ParamType.Weight, self.allowed_weight_range), (ParamType.Bias, self.allowed_bias_range)):
                    param = self._get_param(param_type, node_type, iteration)
                    if param is not None:
                        param.data.clamp_(rng.start, rng.end)

    def _get_param(self, param_type: ParamType, node_type: NodeType, iterative_node_identifier: int):
        return getattr(self, self._param_name(param_type, node_type, iterative_node_identifier), None)

    def _quantize_message(self, x: torch.Tensor, q_bit: int) -> torch.Tensor:
        """torch version of the message quantiser with the straight-through estimator (:187-214)"""
        from .Functions import Functions
        return Functions.cal_msa_q_torch(x, q_bit)

    def fetch_param(self, param_type: ParamType, node_type: NodeType, curr_iter: int) -> Optional[torch.Tensor]:
        sharing_type = self.node_weight_sharing_config.get(node_type)
        if sharing_type in [1, 2, 3]:
            return self._get_param(param_type, node_type, curr_iter)
        if sharing_type in [4, 5]:
            if self.fixed_iterative_nodes and len(self.fixed_iterative_nodes) > 0:
                earlier = [i for i in self.fixed_iterative_nodes if i <= curr_iter]
                pick = max(earlier) if earlier else self.fixed_iterative_nodes[0]
                return self._get_param(param_type, node_type, pick)
            return self._get_param(param_type, node_type, 0)
        return None

    def get_trainable_parameters(self):
        params = []
        for node_type, sharing_type in self.node_weight_sharing_config:
            if sharing_type == 0:
                continue
            if sharing_type in [1, 2, 3]:
                iterations = range(self.iter_node_counts)
            else:
                iterations = self.fixed_iterative_nodes if self.fixed_iterative_nodes else [0]
            for it in iterations:
                if it < self.fixed_iterative_nodes_init_weight:
                    continue
                param = self._get_param(ParamType.Weight, node_type, it)
                if param is not None:
                    params.append(param)
        return params

    # dense buffers of the reference (:85-92) as read-only attributes
    def __getattr__(self, name):
        for key, attr in _BUFFERS:
            if name == key:
                return self.conn_mat.dense(attr, device=self._param_device())
        return super().__getattr__(name)

    def _param_device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    # ---- folding of the sharing types into per-iteration rows ------------------------------------------------
    def _erow(self, device):
        t = self._erow_cache.get(device)
        if t is None:
            t = torch.as_tensor(self.conn_mat.graph.erow, dtype=torch.long, device=device)
            self._erow_cache[device] = t
        return t

    def _edge_row(self, node_type, sharing, w, device):
        """one [E] row from a CN/UCN parameter of sharing type 1-4 (differentiable)"""
        E = int(self.sum_edge)
        w = w.to(device)
        if sharing in (1, 4):
            return w
        if sharing == 2:
            return w.index_select(0, self._erow(device))      # weight per check, expanded by W_skipconn2odd (:455-487)
        if sharing == 3:
            return w.expand(E)
        raise ValueError(f"unsupported {node_type} sharing type {sharing}")

    def fold_weights(self, iterations, device, fixed_iteration=(), fixed_iter_weight=None):
        """-> (vn_w [n,N] | None, cn_w [n,E] | None, ucn_w [n,E] | None, compute_ucn, ucn_mix) for the executed iterations.
        Mirrors the branches of the reference: VN multiplies only for sharing 2/3 (4 with fixed weights; 1 and 5 are no-ops,
        SURVEY.md A.3); CN 0 = no multiply, 1-4; UCN weights are mixed in only when UCN type == CN type in {1,2,3}."""
        cfg = self.node_weight_sharing_config
        cn, ucn, vn = cfg.get(NodeType.CN), cfg.get(NodeType.UCN), cfg.get(NodeType.VN)
        if cn == 5:
            raise ValueError("cn_weight_sharing=5 has no forward branch in the reference (UnboundLocalError there)")
        fixed_iteration = list(fixed_iteration)
        vn_rows, cn_rows, ucn_rows = [], [], []
        fidx = 0
        for t in iterations:
            if vn in (2, 3):
                w = self.fetch_param(ParamType.Weight, NodeType.VN, t).to(device)
                vn_rows.append(w if vn == 2 else w.expand(self.N))
            elif vn == 4:
                w = fixed_iter_weight[fidx] if t in fixed_iteration else self.fetch_param(ParamType.Weight, NodeType.VN, t)
                w = torch.as_tensor(w, dtype=torch.float32, device=device)
                if w.numel() not in (1, self.N):
                    raise ValueError("vn_weight_sharing=4 weights are per edge and cannot scale the [.., N] channel input "
                                     "(the reference raises a shape error here too)")
                vn_rows.append(w.reshape(-1).expand(self.N))
            if cn in (1, 2, 3):
                cn_rows.append(self._edge_row(NodeType.CN, cn, self.fetch_param(ParamType.Weight, NodeType.CN, t), device))
            elif cn == 4:
                w = fixed_iter_weight[fidx] if t in fixed_iteration else self.fetch_param(ParamType.Weight, NodeType.CN, t)
                cn_rows.append(torch.as_tensor(w, dtype=torch.float32, device=device).reshape(-1))
            if ucn == cn and cn in (1, 2, 3):
                ucn_rows.append(self._edge_row(NodeType.UCN, ucn, self.fetch_param(ParamType.Weight, NodeType.UCN, t), device))
            if t in fixed_iteration:
                fidx += 1
        stack = lambda rows: torch.stack(rows).contiguous() if rows else None   # noqa: E731
        return stack(vn_rows), stack(cn_rows), stack(ucn_rows), ucn > 0, bool(ucn_rows)

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(
            self,
            xa: torch.Tensor | list[torch.Tensor],
            target_iter: int | list[int] = None,
            fixed_iter: int | list[int] = None,
            fixed_iter_weight: torch.Tensor | list[torch.Tensor] = None,
    ):
        """Same contract as the reference (:260-284): xa [batch, N, Z] (or a list, one per executed iteration);
        target_iter None / int / list; fixed_iter + fixed_iter_weight for sharing type 4."""
        if isinstance(target_iter, int):
            iteration = [target_iter]
        elif isinstance(target_iter, list):
            iteration = target_iter
        else:
            iteration = list(range(self.iter_node_counts))
        if fixed_iter is not None:
            for each_iter in fixed_iter:           # an int is not iterable: TypeError, as in the reference (:294)
                if each_iter not in iteration:
                    iteration.append(each_iter)
        iteration = sorted(iteration)

        is_input_iterable = isinstance(xa, list)
        if is_input_iterable:
            assert len(xa) == len(iteration) - len(fixed_iter)
            assert isinstance(xa[0], torch.Tensor)
        fixed_iteration = fixed_iter if isinstance(fixed_iter, list) else ([fixed_iter] if isinstance(fixed_iter, int) else [])
        if len(fixed_iteration) > 0:
            assert len(fixed_iteration) == len(fixed_iter_weight)

        first = xa[iteration[0]] if is_input_iterable else xa
        if tuple(first.shape) != (self.batch_size, self.N, self.Z):
            raise RuntimeError(f"input shape {tuple(first.shape)} does not match (batch_size, N, Z) = "
                               f"{(self.batch_size, self.N, self.Z)} given to the constructor")
        device = first.device
        gid = self.conn_mat.graph_id(device)

        # maximal runs of consecutive iterations -> one launch each (list-xa: one iteration per launch)
        if self.store_llr not in ("all", "last", "none"):
            raise ValueError(f"store_llr must be 'all', 'last' or 'none', got {self.store_llr!r}")
        runs, cur = [], []
        for t in iteration:
            if cur and (t != cur[-1] + 1 or is_input_iterable):
                runs.append(cur)
                cur = []
            cur.append(t)
        if cur:
            runs.append(cur)

        cfg = self.node_weight_sharing_config
        dec = {DecoderType.SP: 0, DecoderType.MS: 1, DecoderType.QMS: 2}[self.decoding_type]
        xin_state = None
        done_fixed = 0
        for run in runs:
            t0, t1 = run[0], run[-1]
            n_fixed = sum(1 for t in run if t in fixed_iteration)
            fw = None if fixed_iter_weight is None else list(fixed_iter_weight)[done_fixed:done_fixed + n_fixed]
            inference = not torch.is_grad_enabled()
            if inference and not fixed_iteration and t0 == 0:
                # validation loops under no_grad: live gather from the flat parameter vector instead of ~2 T folding launches
                vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = _folded_live(self, len(run), device)
            else:
                vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = self.fold_weights(run, device, fixed_iteration, fw)
            done_fixed += n_fixed
            x_run = xa[t0] if is_input_iterable else xa
            if is_input_iterable:
                xin_state = None                     # xa_input is re-assigned from the iteration's own input (:321-323)
            llr_init = None
            if t0 > 0:
                if not self._llr_valid[t0]:
                    raise RuntimeError(f"iteration {t0} continues from self.llr[{t0}], but the call that last executed iteration "
                                       f"{t0 - 1} did not store it (model.store_llr = {self.store_llr!r}); use store_llr = 'all'")
                llr_init = self.llr[t0].detach().to(device)            # a constant of this call, as a graph-less tensor is in the reference
            app_init = self.outputs[t0 - 1].detach().to(device) if (compute_ucn and t0 > 0) else None
            llr_mode = {"none": 0, "last": 1, "all": 2}[self.store_llr]
            want_xin = len(runs) > 1 and not is_input_iterable
            needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (vn_w, cn_w, ucn_w))
            # Two configurations have no backward on this path: the SP decoder (the sweep kernels cover MS / QMS) and a
            # target_iter list with gaps when VN weights are present (the compounding channel-input chain would have to carry
            # gradients across launches).  Their forward works as in the reference; loss.backward() raises, and a warning
            # says so now rather than only then.
            no_bwd = self.decoding_type == DecoderType.SP or (xin_state is not None and vn_w is not None)
            if needs_grad and no_bwd and not self.__dict__.get("_warned_no_bwd"):
                import warnings
                warnings.warn("BoostedNeuralLDPCDecoder (B200): this call is forward-only — "
                              + ("the SP decoder has no backward kernel" if self.decoding_type == DecoderType.SP else
                                 "a target_iter list with gaps cannot be trained when VN weights are present")
                              + "; backward() through its outputs will raise.  Use torch.no_grad() for inference.")
                self.__dict__["_warned_no_bwd"] = True
            want_dump = needs_grad and not no_bwd
            run_op = ops.boosted_forward_direct if inference else torch.ops.nldpc.boosted_forward      # no autograd state: skip the dispatcher
            soft, llr_out, xin_out, _, _ = run_op(
                x_run, vn_w, cn_w, ucn_w, gid, len(run), dec, int(self.decoder_qms_qbit),
                float(self.allowed_llr_range.start), float(self.allowed_llr_range.end), bool(compute_ucn), bool(ucn_mix),
                llr_init, xin_state, app_init, llr_mode, want_xin, 1, 0, want_dump,
                # 16-byte state rows (BG2: E = 197 at pitch 200) let the kernels export self.llr with vector stores; the
                # tensors stored below are the [..., :E] views (same shape and values as the reference's, :512)
                **({"pad_llr": True} if inference else {}))
            for k, t in enumerate(run):
                self.outputs[t] = soft[k]
                if llr_mode == 2:
                    self.llr[t + 1] = llr_out[k].detach()
                self._llr_valid[t + 1] = llr_mode == 2
            if llr_mode == 1:
                self.llr[t1 + 1] = llr_out.detach()
                self._llr_valid[t1 + 1] = True
            if want_xin:
                xin_state = xin_out.detach()

        if isinstance(target_iter, int):
            return self.outputs[target_iter]
        elif isinstance(target_iter, list):
            return [self.outputs[i] for i in target_iter]
        return self.outputs


def _fused_bce_loss(self, xa, y, etha=1.0, coeff_param=None):
    """`LDPCDecoderLoss(LossType.BCE, etha)(self(xa), y, coeff_param)` (train/train_BoostedNeuralLDPCDecoder.py:278-289) as ONE
    forward launch — the T iterations, the loss and dL/dout without any [T, B, N*Z] tensor crossing HBM twice — whose backward
    is the sweep kernel (ops.boosted_train_loss).  Differentiable w.r.t. the parameters like the two-call form.  Returns None
    when the configuration is not covered (other codes, SP, UCN weights, no CN weights, sharing type 4 / 5, list inputs); the
    caller then uses the two-call form.  Unlike forward() it does not refresh self.outputs / self.llr; labels must be 0 / 1."""
    if not isinstance(xa, torch.Tensor) or not xa.is_cuda or tuple(xa.shape) != (self.batch_size, self.N, self.Z):
        return None
    cfg = self.node_weight_sharing_config
    cn, ucn, vn = cfg.get(NodeType.CN), cfg.get(NodeType.UCN), cfg.get(NodeType.VN)
    if self.decoding_type not in (DecoderType.MS, DecoderType.QMS) or ucn != 0 or cn not in (1, 2, 3) or vn not in (0, 1, 2, 3, 5):
        return None
    T, device = int(self.iter_node_counts), xa.device
    dec = 1 if self.decoding_type == DecoderType.MS else 2
    gid = self.conn_mat.graph_id(device)
    lo, hi = float(self.allowed_llr_range.start), float(self.allowed_llr_range.end)
    if not ops.boosted_train_covered(gid, dec, int(self.decoder_qms_qbit), lo, hi, T, True, vn in (2, 3)):
        return None
    vn_w, cn_w, _, _, _ = self.fold_weights(list(range(T)), device)
    coef = ops.iteration_coefs(T, etha, coeff_param, device)
    return 1.0 * ops.boosted_train_loss(xa, vn_w, cn_w, ops.pack_labels(y.to(torch.float32)), coef, gid, T, dec,
                                        int(self.decoder_qms_qbit), lo, hi)


BoostedNeuralLDPCDecoder.fused_bce_loss = _fused_bce_loss


def _folded_live(self, T, device):
    """fold_weights(range(T)) on `device` for the decode-only paths, read LIVE from the parameters: the parameters are slices
    of one flat vector (_flatparams), so every folded row set is ONE gather through an index map (<= 3 launches instead of ~2 T
    expand / stack launches).  Only the index maps are cached, keyed by the parameters' data pointers — never a value, so
    `p.data.clamp_()` (the reference's `_apply_constraints`) or a kernel writing the weights cannot leave a stale copy behind.
    Sharing types 4 / 5, parameters spread over several storages (`p.data = other`) or another device: plain fold_weights."""
    from .._flatparams import element_offset, storage_base
    cfg = self.node_weight_sharing_config
    cn, ucn, vn = cfg.get(NodeType.CN), cfg.get(NodeType.UCN), cfg.get(NodeType.VN)
    ps = list(self.parameters())
    key = (device, T, tuple(p.data_ptr() for p in ps))
    cache = self.__dict__.setdefault("_fold_index", {})
    hit = cache.get((device, T))
    if hit is None or hit[0] != key:
        maps = None
        base = storage_base(ps, device) if (cn in (0, 1, 2, 3) and vn in (0, 1, 2, 3, 5) and ucn in (0, 1, 2, 3) and ps) else None
        if base is not None:
            E, N = int(self.sum_edge), self.N
            erow = self._erow(device)
            ar_e, ar_n = torch.arange(E, device=device), torch.arange(N, device=device)

            def index_rows(node_type, sharing, cols):
                rows = []
                for t in range(T):
                    off = element_offset(self.fetch_param(ParamType.Weight, node_type, t), base)
                    if cols == N:
                        rows.append(off + ar_n if sharing == 2 else torch.full((N,), off, dtype=torch.long, device=device))
                    else:
                        rows.append(off + ar_e if sharing == 1 else (off + erow if sharing == 2
                                                                      else torch.full((E,), off, dtype=torch.long, device=device)))
                return torch.stack(rows).contiguous()

            maps = (index_rows(NodeType.VN, vn, N) if vn in (2, 3) else None,
                    index_rows(NodeType.CN, cn, E) if cn in (1, 2, 3) else None,
                    index_rows(NodeType.UCN, ucn, E) if (ucn == cn and cn in (1, 2, 3)) else None, base)
        hit = (key, maps)
        if len(cache) > 8:
            cache.clear()
        cache[(device, T)] = hit
    if hit[1] is None:
        with torch.no_grad():
            folded = self.fold_weights(list(range(T)), device)
        return tuple(t.detach() if isinstance(t, torch.Tensor) else t for t in folded)
    i_vn, i_cn, i_ucn, base = hit[1]
    take = lambda idx: None if idx is None else base[idx]       # noqa: E731
    return take(i_vn), take(i_cn), take(i_ucn), ucn > 0, i_ucn is not None


def _decode(self, xa, n_iters, soft_mode, hard_mode):
    from .. import ops
    T = self.iter_node_counts if n_iters is None else n_iters
    device = xa.device
    gid = self.conn_mat.graph_id(device)
    dec = {DecoderType.SP: 0, DecoderType.MS: 1, DecoderType.QMS: 2}[self.decoding_type]
    tail = (gid, T, dec, int(self.decoder_qms_qbit), float(self.allowed_llr_range.start), float(self.allowed_llr_range.end))
    # decode-only: the folded [T, .] weight rows are gathered live from the flat parameter vector (_folded_live; captured into
    # a CUDA graph the gathers belong to the graph, so a replay re-reads the parameters); no dispatcher
    vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = _folded_live(self, T, device)
    soft, _, _, hard, _ = ops.boosted_forward_direct(xa, vn_w, cn_w, ucn_w, *tail, bool(compute_ucn), bool(ucn_mix), None, None, None,
                                                     0, False, soft_mode, hard_mode, False)
    return soft, hard


@torch.no_grad()
def decode_hard(self, xa, n_iters=None, all_iters=False):
    """Stateless throughput mode (any batch size): packed hard decisions `(out < 0)` (Functions.py:90 predicate) after
    `n_iters` iterations from the zero state; uint8 [B, ceil(N*Z/8)] or [T, B, ...] with all_iters."""
    return _decode(self, xa, n_iters, 0, 1 if all_iters else 2)[1]


@torch.no_grad()
def decode_soft_last(self, xa, n_iters=None):
    """Stateless: the soft output of the last iteration only, [B, N*Z]."""
    return _decode(self, xa, n_iters, 2, 0)[0]


@torch.no_grad()
def decode_host_q8(self, xq_cpu, scale=0.5, device=None, n_iters=None, soft=False, hard=True):
    """End-to-end host API with one-byte channel LLRs: `xq_cpu` int8 CPU tensor [B, N, Z], x = scale * q (the Boosted
    pipeline's inputs are quantised by Functions.Cal_MSA_Q before they reach the decoder — q_bit 5: multiples of 0.5 in
    +-7.5 — so q = round(x / scale) carries them without loss at a quarter of the PCIe bytes).  Stateless decode from the
    zero state; returns (soft [B, N*Z] of the last iteration | None, packed hard decisions [B, ceil(N*Z/8)] | None) on the host."""
    from .. import _lib, ops
    device = torch.device(device if device is not None else self._param_device())
    if device.type != "cuda":
        device = torch.device("cuda")
    T = self.iter_node_counts if n_iters is None else n_iters
    gid = self.conn_mat.graph_id(device)
    vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = _folded_live(self, T, self._param_device())
    host = lambda t: None if t is None else t.cpu().contiguous()       # noqa: E731
    dec = {DecoderType.SP: 0, DecoderType.MS: 1, DecoderType.QMS: 2}[self.decoding_type]
    return ops.boosted_decode_host_q8(gid, xq_cpu, scale, host(vn_w), host(cn_w), host(ucn_w), T, dec, int(self.decoder_qms_qbit),
                                      float(self.allowed_llr_range.start), float(self.allowed_llr_range.end), bool(compute_ucn),
                                      bool(ucn_mix), _lib.NLDPC_OUT_LAST if soft else _lib.NLDPC_OUT_NONE,
                                      _lib.NLDPC_OUT_LAST if hard else _lib.NLDPC_OUT_NONE)


BoostedNeuralLDPCDecoder.decode_hard = decode_hard
BoostedNeuralLDPCDecoder.decode_soft_last = decode_soft_last
BoostedNeuralLDPCDecoder.decode_host_q8 = decode_host_q8


def _add_dense_buffers(module, state_dict, prefix, local_metadata):
    """state_dict hook: the reference's buffer keys after the parameters (reference order)"""
    dev = module._param_device()
    for key, attr in _BUFFERS:
        state_dict[prefix + key] = module.conn_mat.dense(attr, device=dev)
    return state_dict


def _drop_dense_buffers(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """load_state_dict pre-hook: the dense structure matrices are derived from the base graph, so they are not loaded — but a
    checkpoint written for ANOTHER graph must fail like the reference's strict load does (size mismatch) instead of silently
    attaching its weights to this graph: every buffer present in the checkpoint is compared with the synthesised one."""
    for key, attr in _BUFFERS:
        t = state_dict.pop(prefix + key, None)
        if t is None:
            if strict:
                missing_keys.append(prefix + key)
            continue
        own = module.conn_mat.dense(attr, device=torch.device("cpu"))
        if tuple(t.shape) != tuple(own.shape):
            error_msgs.append(f"size mismatch for {prefix + key}: copying a param with shape {tuple(t.shape)} from checkpoint, "
                              f"the shape in current model is {tuple(own.shape)}.")
        elif not torch.equal(t.detach().to("cpu", own.dtype), own):
            error_msgs.append(f"{prefix + key} in the checkpoint differs from the matrix this module derives from its base graph "
                              "(the checkpoint was written for a different code)")
