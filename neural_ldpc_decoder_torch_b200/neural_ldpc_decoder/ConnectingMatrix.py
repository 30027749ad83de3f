"""ConnectingMatrix — same constructor and attributes as the reference class
(/root/reference/src/neural_ldpc_decoder/ConnectingMatrix.py:3-140), built from the sparse
edge tables of `graph.TannerGraph` instead of Python loops over dense matrices.

The dense 0/1 matrices (`W_odd2even`, `W_skipconn2even`, `W_even2odd`, `W_output`,
`lifting_matrix_1/2`) are only needed for `state_dict()` / checkpoint compatibility — the B200
decode path never reads them — so they are materialised lazily on first access.
"""
import numpy as np

from ..graph import TannerGraph


class ConnectingMatrix:
    _DENSE = ("W_odd2even", "W_skipconn2even", "W_even2odd", "W_output", "lifting_matrix_1", "lifting_matrix_2")

    def __init__(
            self,
            Z: int,
            basegraph: np.ndarray,
            dtype_w_odd2even=np.float32,
            dtype_w_skipconn2even=np.float32,
            dtype_w_even2odd=np.float32,
            dtype_w_output=np.float32,
            dtype_lifting_matrix=np.float32
    ):
        self.graph = TannerGraph(basegraph, Z)
        self.basegraph = np.asarray(basegraph).copy()
        self.M, self.N = self.basegraph.shape
        self.Z = Z
        self.basegraph_binary = np.where(self.basegraph == -1, 0, 1).astype(self.basegraph.dtype)
        self.sum_edge_c = np.sum(self.basegraph_binary, axis=1)
        self.sum_edge_v = np.sum(self.basegraph_binary, axis=0)
        self.sum_edge = np.sum(self.sum_edge_v)
        self.dtype_w_odd2even = dtype_w_odd2even
        self.dtype_w_skipconn2even = dtype_w_skipconn2even
        self.dtype_w_even2odd = dtype_w_even2odd
        self.dtype_w_output = dtype_w_output
        self.dtype_lifting_matrix = dtype_lifting_matrix
        self.neurons_per_even_layer = np.copy(self.sum_edge)
        self.neurons_per_odd_layer = np.copy(self.sum_edge)
        self._dense = {}

    # ---- lazy dense views (reference layout: "even" layer = column-major edges, "odd" layer = row-major) ----
    def __getattr__(self, name):
        if name in type(self)._DENSE:
            d = self.__dict__.setdefault("_dense", {})
            if name not in d:
                d[name] = self._build_dense(name)
            return d[name]
        raise AttributeError(name)

    def _build_dense(self, name):
        g = self.graph
        E, Z = g.E, g.Z
        rm = np.arange(E)
        cm_of = g.rm_to_cm                     # rm -> cm
        if name == "W_skipconn2even":          # [N, E(cm)]: channel LLR of block j feeds its edges (:135-140)
            W = np.zeros((g.N, E), dtype=self.dtype_w_skipconn2even)
            W[g.ecol, cm_of] = 1.0
            return W
        if name == "W_output":                 # [E(rm), N]  (:122-132)
            W = np.zeros((E, g.N), dtype=self.dtype_w_output)
            W[rm, g.ecol] = 1.0
            return W
        if name == "W_odd2even":               # [E(rm), E(cm)]: other edges of the same variable block (:87-105)
            W = (g.ecol[:, None] == g.ecol[None, :]) & (rm[:, None] != rm[None, :])   # [rm', rm]
            out = np.zeros((E, E), dtype=self.dtype_w_odd2even)
            out[:, cm_of] = W.astype(self.dtype_w_odd2even)
            return out
        if name in ("W_even2odd", "W_even2odd_with_self"):   # [E(cm), E(rm)]: edges of the same check (:107-120)
            W = (g.erow[:, None] == g.erow[None, :])
            if name == "W_even2odd":
                W = W & (rm[:, None] != rm[None, :])
            out = np.zeros((E, E), dtype=self.dtype_w_even2odd)
            out[cm_of, :] = W.astype(self.dtype_w_even2odd)
            return out
        if name == "W_skipconn2odd":           # [M, E(rm)] (boosted ConnectingMatrix.py:157-163)
            W = np.zeros((g.M, E), dtype=np.float32)
            W[g.erow, rm] = 1.0
            return W
        if name in ("lifting_matrix_1", "lifting_matrix_2"):
            # ones at (k*Z + h, k*Z + (h + s) % Z); k = column-major index for matrix 1, row-major for matrix 2 (:69-85)
            L = np.zeros((E * Z, E * Z), dtype=self.dtype_lifting_matrix)
            k = cm_of if name == "lifting_matrix_1" else rm
            h = np.arange(Z)
            rows = (k[:, None] * Z + h[None, :]).reshape(-1)
            cols = (k[:, None] * Z + (h[None, :] + g.eshift[:, None]) % Z).reshape(-1)
            L[rows, cols] = 1
            return L
        raise AttributeError(name)
