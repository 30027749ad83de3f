"""ConnectingMatrixTorch — same constructor/attributes as the reference
(/root/reference/src/neural_ldpc_decoder/ConnectingMatrixTorch.py:6-46).  The torch copies of the dense
matrices are lazy (checkpoint compatibility only); the hot path uses `graph_id(device)` — a handle to the
edge/shift tables that live on the GPU (nldpc_graph_create)."""
import numpy as np
import torch

from .ConnectingMatrix import ConnectingMatrix

_DENSE_DTYPE = {
    "W_odd2even": "dtype_w_odd2even", "W_skipconn2even": "dtype_w_skipconn2even", "W_even2odd": "dtype_w_even2odd",
    "W_output": "dtype_w_output", "lifting_matrix_1": "dtype_lifting_matrix", "lifting_matrix_2": "dtype_lifting_matrix",
}


class ConnectingMatrixTorch:
    _DENSE_DTYPE = _DENSE_DTYPE

    def __init__(self, connecting_matrix: ConnectingMatrix, device: torch.device = torch.device("cpu"),
                 dtype_w_odd2even: torch.dtype = torch.float32, dtype_w_skipconn2even: torch.dtype = torch.float32,
                 dtype_w_even2odd: torch.dtype = torch.float32, dtype_w_output: torch.dtype = torch.float32,
                 dtype_lifting_matrix: torch.dtype = torch.float32):
        cm = self._cm = connecting_matrix
        self.device, self.graph = torch.device(device), cm.graph
        self.N, self.M, self.Z = cm.N, cm.M, cm.Z
        for name in ("basegraph", "sum_edge_c", "sum_edge_v", "sum_edge"):          # own copies, as the reference keeps
            setattr(self, name, getattr(cm, name).copy())
        self.neurons_per_even_layer, self.neurons_per_odd_layer = np.copy(self.sum_edge), np.copy(self.sum_edge)
        self.dtype_w_odd2even, self.dtype_w_skipconn2even = dtype_w_odd2even, dtype_w_skipconn2even
        self.dtype_w_even2odd, self.dtype_w_output = dtype_w_even2odd, dtype_w_output
        self.dtype_lifting_matrix = dtype_lifting_matrix
        self._dense_t = {}      # dense matrix name -> tensor on self.device, filled on first access

    def __getattr__(self, name):
        table = type(self)._DENSE_DTYPE
        if name in table:
            d = self.__dict__.setdefault("_dense_t", {})
            if name not in d:
                d[name] = torch.tensor(getattr(self._cm, name), dtype=getattr(self, table[name]), device=self.device)
            return d[name]
        raise AttributeError(name)

    def dense(self, name, device=None):
        """Dense matrix `name` as a tensor on `device` (default: this object's device), not cached."""
        t = torch.as_tensor(getattr(self._cm, name))
        return t.to(device=device if device is not None else self.device, dtype=getattr(self, type(self)._DENSE_DTYPE[name]))

    def graph_id(self, device):
        """Handle of the device-resident Tanner tables for CUDA device `device` (created on first use)."""
        from .. import _lib
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.NldpcError(f"the B200 decode path needs a CUDA device, got {device}")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        cache = self.__dict__.setdefault("_graph_ids", {})      # per device: skips hashing the base graph on every call
        gid = cache.get(idx)
        if gid is None:
            gid = cache[idx] = _lib.graph_id_for(self.basegraph, self.Z, idx)
        return gid
