"""NeuralLDPCDecoder — drop-in for /root/reference/src/neural_ldpc_decoder/NeuralLDPCDecoder.py:6-100.

Same constructor, attributes, parameter names (`weights_var.{t}`, `biases_var.{t}`, row-major edge order,
init 0.5 / 0), forward signature and return value (list of T tensors [B, N*Z], differentiable w.r.t. the
parameters).  The T-iteration loop (:54-98) is ONE sm_100a kernel behind `torch.ops.nldpc.neural_forward`.

`state_dict()` carries the reference's dense buffers (`W_odd2even`, …, `Lift_Matrix2`) so checkpoints
interchange with the reference (CheckPointUtil.load is strict, CheckPointUtil.py:154); they are synthesised
on demand and never occupy GPU memory.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .ConnectingMatrixTorch import ConnectingMatrixTorch
from .. import ops  # noqa: F401  (registers torch.ops.nldpc.*)

_BUFFERS = (("W_odd2even", "W_odd2even"), ("W_skipconn2even", "W_skipconn2even"), ("W_even2odd", "W_even2odd"),
            ("W_output", "W_output"), ("Lift_Matrix1", "lifting_matrix_1"), ("Lift_Matrix2", "lifting_matrix_2"))


class NeuralLDPCDecoder(nn.Module):
    def __init__(
            self,
            iter_node_counts,
            batch_size,
            connecting_matrix: ConnectingMatrixTorch,
    ):
        super(NeuralLDPCDecoder, self).__init__()
        self.iter_node_counts = iter_node_counts
        self.batch_size = batch_size      # stored, unused in forward (reference :45 uses xa.shape[0])

        self.conn_mat = connecting_matrix

        self.N = self.conn_mat.N
        self.M = self.conn_mat.M
        self.Z = self.conn_mat.Z
        self.sum_edge = self.conn_mat.sum_edge
        self.neurons_per_odd_layer = self.conn_mat.neurons_per_odd_layer
        self.neurons_per_even_layer = self.conn_mat.neurons_per_even_layer

        self.weights_var = nn.ParameterList([
            nn.Parameter(0.5 * torch.ones(int(self.conn_mat.sum_edge), dtype=torch.float32))
            for _ in range(iter_node_counts)
        ])
        self.biases_var = nn.ParameterList([
            nn.Parameter(torch.zeros(int(self.conn_mat.sum_edge), dtype=torch.float32))
            for _ in range(iter_node_counts)
        ])
        self._register_state_dict_hook(_add_dense_buffers)
        self.register_load_state_dict_pre_hook(_drop_dense_buffers)
        self._flatten_params()

    # dense buffers of the reference (:27-32) as read-only attributes, for code that inspects them
    def __getattr__(self, name):
        for key, attr in _BUFFERS:
            if name == key:
                return self.conn_mat.dense(attr, device=self._param_device())
        return super().__getattr__(name)

    def _param_device(self):
        return self.weights_var[0].device if len(self.weights_var) else torch.device("cpu")

    def _stacked(self):
        return torch.stack(list(self.weights_var)), torch.stack(list(self.biases_var))

    def _flatten_params(self):
        from .._flatparams import flatten_
        flatten_(list(self.weights_var) + list(self.biases_var))       # weights [T, E] then biases [T, E], consecutive rows

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)                      # .to() / .cuda() give every parameter its own storage again
        self._flatten_params()
        return out

    def _stacked_nograd(self, device):
        """[T, E] weights / biases on `device` for decode-only calls WITHOUT a stack launch: the parameters are consecutive rows
        of one flat vector (see _flatparams), so the rows are a strided view — live, never a stale copy.  Only the layout check
        is cached (keyed by the data pointers); a broken layout (`p.data = other`) or another device falls back to stacking."""
        ws, bs = list(self.weights_var), list(self.biases_var)
        ptrs = tuple(p.data_ptr() for p in ws) + tuple(p.data_ptr() for p in bs)
        hit = self.__dict__.get("_rows_view")
        if hit is None or hit[0] != ptrs or hit[1] != device:
            views = None
            T, E = len(ws), ws[0].numel()

            def rows(ps):
                p0 = ps[0]
                sp = p0.untyped_storage().data_ptr()
                ok = all(p.device == device and p.is_contiguous() and p.dtype == torch.float32 and p.untyped_storage().data_ptr() == sp
                         and p.data_ptr() == p0.data_ptr() + 4 * E * t for t, p in enumerate(ps))
                return torch.as_strided(p0.detach(), (T, E), (E, 1)) if ok else None

            w, b = rows(ws), rows(bs)
            if w is not None and b is not None:
                views = (w, b)
            hit = (ptrs, device, views)
            self.__dict__["_rows_view"] = hit
        if hit[2] is not None:
            return hit[2]
        with torch.no_grad():
            w, b = self._stacked()
        return w.detach().to(device), b.detach().to(device)

    def forward(self, xa):
        """xa [B, N, Z] float32 on a CUDA device -> list of T tensors [B, N*Z] (iteration outputs, :94-98)."""
        gid = self.conn_mat.graph_id(xa.device)
        if not torch.is_grad_enabled():
            # inference: live [T, E] views of the flat parameter vector, no stack launches, no dispatcher (see _stacked_nograd)
            return list(ops.neural_forward_direct(xa, *self._stacked_nograd(xa.device), gid).unbind(0))
        w, b = self._stacked()
        if w.device != xa.device:
            w, b = w.to(xa.device), b.to(xa.device)
        if torch.is_grad_enabled() and (w.requires_grad or b.requires_grad):
            # training: the forward also spills the per-iteration v2c the backward kernel needs (no forward re-run)
            out, _ = torch.ops.nldpc.neural_forward_train(xa, w, b, gid)
        else:
            out = torch.ops.nldpc.neural_forward(xa, w, b, gid)
        return list(out.unbind(0))

    @torch.no_grad()
    def decode_hard(self, xa, all_iters=False):
        """Throughput mode: packed hard decisions `(out < 0)` (Functions.py:90 predicate), uint8
        [B, ceil(N*Z/8)] of the last iteration (or [T, B, ...]); soft outputs are never written to HBM."""
        gid = self.conn_mat.graph_id(xa.device)
        w, b = self._stacked_nograd(xa.device)           # live views (or a fresh stack): a captured launch re-reads the parameters on replay
        return ops.neural_hard_direct(xa, w, b, gid, all_iters)

    @torch.no_grad()
    def decode_host(self, xa_cpu, device=None, soft=False, hard=True, scale=1.0):
        """End-to-end host API: CPU tensor in, CPU results out (chunked H2D / decode / D2H overlap).  xa_cpu float32, or
        float16 values / int8 codes (x = scale * q) for callers whose LLRs are quantised anyway: the link carries 2 / 1 byte per
        LLR instead of 4 and the result equals the decode of the widened values bit for bit."""
        from .. import _lib
        device = torch.device(device if device is not None else "cuda")
        gid = self.conn_mat.graph_id(device)
        w, b = self._stacked_nograd(self._param_device())
        return ops.neural_decode_host(gid, xa_cpu.contiguous(), w.cpu().contiguous(), b.cpu().contiguous(),
                                      _lib.NLDPC_OUT_ALL if soft else _lib.NLDPC_OUT_NONE,
                                      _lib.NLDPC_OUT_LAST if hard else _lib.NLDPC_OUT_NONE, scale=scale)


def _add_dense_buffers(module, state_dict, prefix, local_metadata):
    """state_dict hook: emit the reference's buffer keys, in the reference's order (buffers first)."""
    dev = module._param_device()
    items = list(state_dict.items())
    own = [(k, v) for k, v in items if k.startswith(prefix)]
    for k, _ in own:
        del state_dict[k]
    for key, attr in _BUFFERS:
        state_dict[prefix + key] = module.conn_mat.dense(attr, device=dev)
    for k, v in own:
        state_dict[k] = v
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
