"""Tanner-graph tables of a quasi-cyclic (protograph) LDPC code.

The reference materialises the graph as dense 0/1 matrices
(/root/reference/src/neural_ldpc_decoder/ConnectingMatrix.py:68-140, boosted variant :82-163).
Here the same structure is an edge list: everything the CUDA path needs is derived from
`basegraph` (int matrix, -1 = no edge, entry = circulant shift, used modulo Z) and `Z`.

Edge order conventions (they are the parameter/`state_dict` layout of the reference):
  * row-major index rm(e) (check i outer, variable j inner)  — `weights_var[t][rm]`,
    `biases_var[t][rm]`, per-edge boosted weights, `llr[:, :, rm]`   (ConnectingMatrix.py:78-85)
  * column-major index cm(e) (variable j outer, check i inner) — rows of `W_even2odd`,
    columns of `W_skipconn2even`/`W_odd2even`                        (ConnectingMatrix.py:69-76)
"""
import json
import os

import numpy as np

_RES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")

BUILTIN_GRAPHS = {"nr_bg2_set0": "nr_bg2_set0.json", "wimax_n576_r34": "wimax_n576_r34.json"}


def load_basegraph(name):
    """Return (basegraph [M,N] int64 with -1 for 'no edge', default Z) of a built-in code:
    'nr_bg2_set0' (5G NR BG2, set 0; reference resources/basegraph2_set0.txt) or
    'wimax_n576_r34' (802.16e N=576 R=3/4; reference resources/wman_N0576_R34_z24.txt)."""
    with open(os.path.join(_RES, BUILTIN_GRAPHS[name])) as f:
        doc = json.load(f)
    bg = -np.ones((doc["M"], doc["N"]), dtype=np.int64)
    for i, row in enumerate(doc["rows"]):
        for j, s in row:
            bg[i, j] = s
    return bg, int(doc["Z_default"])


class TannerGraph:
    """Sparse edge tables of basegraph lifted by Z (no dense matrices)."""

    def __init__(self, basegraph, Z):
        bg = np.asarray(basegraph)
        if bg.ndim != 2:
            raise ValueError("basegraph must be a 2-D integer matrix")
        self.basegraph = bg.astype(np.int64).copy()
        self.M, self.N = (int(v) for v in bg.shape)
        self.Z = int(Z)
        if self.Z < 1:
            raise ValueError("Z must be a positive integer")
        ii, jj = np.nonzero(self.basegraph != -1)          # row-major enumeration
        self.erow = ii.astype(np.int32)
        self.ecol = jj.astype(np.int32)
        self.eshift = (self.basegraph[ii, jj] % self.Z).astype(np.int32)
        self.E = int(ii.size)
        self.row_deg = np.bincount(self.erow, minlength=self.M).astype(np.int32)
        self.col_deg = np.bincount(self.ecol, minlength=self.N).astype(np.int32)
        self.row_ptr = np.concatenate([[0], np.cumsum(self.row_deg)]).astype(np.int32)
        # column lists: rm indices of each column's edges in ascending check row
        order = np.lexsort((self.erow, self.ecol))          # sort by col, then row == column-major order
        self.cm_to_rm = order.astype(np.int32)              # cm index -> rm index
        self.rm_to_cm = np.empty(self.E, np.int32)
        self.rm_to_cm[order] = np.arange(self.E, dtype=np.int32)
        self.col_ptr = np.concatenate([[0], np.cumsum(self.col_deg)]).astype(np.int32)
        self.col_edges = self.cm_to_rm                      # [E] grouped by column

    @property
    def key(self):
        return (self.M, self.N, self.Z, self.basegraph.tobytes())

    def basegraph_i32(self):
        return np.ascontiguousarray(self.basegraph, dtype=np.int32)

    # ---- lifted parity-check matrix / systematic encoder (used by tests and the data generator) ----
    def lifted_H(self):
        """Dense [M*Z, N*Z] uint8 H.  Circulant convention (verified against the reference's
        lifting matrices): base entry s has ones at (row h, col (h + s) mod Z)."""
        Z = self.Z
        H = np.zeros((self.M * Z, self.N * Z), dtype=np.uint8)
        h = np.arange(Z)
        for i, j, s in zip(self.erow, self.ecol, self.eshift):
            H[i * Z + h, j * Z + (h + s) % Z] = 1
        return H

    def systematic_generator(self):
        """[K*Z, N*Z] uint8 systematic generator G = [I | P] with H G^T = 0 over GF(2)
        (information bits = first K*Z codeword positions).  Raises if the parity part of H is singular."""
        H = self.lifted_H()
        mz, nz = H.shape
        kz = nz - mz
        A = H[:, :kz].copy()
        Bm = H[:, kz:].copy()
        aug = np.concatenate([Bm, A], axis=1).astype(np.uint8)   # solve Bm * P^T = A
        r = 0
        for c in range(mz):
            piv = np.nonzero(aug[r:, c])[0]
            if piv.size == 0:
                raise ValueError("parity part of H is singular: no systematic generator with K leading info bits")
            p = r + piv[0]
            if p != r:
                aug[[r, p]] = aug[[p, r]]
            rows = np.nonzero(aug[:, c])[0]
            rows = rows[rows != r]
            aug[rows] ^= aug[r]
            r += 1
        Pt = aug[:, mz:]                                         # [mz, kz]
        G = np.concatenate([np.eye(kz, dtype=np.uint8), Pt.T], axis=1)
        return G
