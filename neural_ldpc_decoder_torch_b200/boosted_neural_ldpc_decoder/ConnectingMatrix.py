"""Boosted ConnectingMatrix: the neural one plus `W_even2odd_with_self` and `W_skipconn2odd`
(reference: src/boosted_neural_ldpc_decoder/ConnectingMatrix.py:4-163)."""
import numpy as np

from ..neural_ldpc_decoder.ConnectingMatrix import ConnectingMatrix as _Base


class ConnectingMatrix(_Base):
    _DENSE = _Base._DENSE + ("W_even2odd_with_self", "W_skipconn2odd")

    def __init__(
            self,
            Z: int,
            basegraph: np.ndarray,
            dtype_w_odd2even=np.float32,
            dtype_w_skipconn2even=np.float32,
            dtype_w_even2odd=np.float32,
            dtype_w_even2odd_with_self=np.float32,
            dtype_w_output=np.float32,
            dtype_w_skipconn2odd=np.float32,
            dtype_lifting_matrix=np.float32
    ):
        super().__init__(Z, basegraph, dtype_w_odd2even, dtype_w_skipconn2even, dtype_w_even2odd, dtype_w_output,
                         dtype_lifting_matrix)
        self.dtype_w_even2odd_with_self = dtype_w_even2odd_with_self
        self.dtype_w_skipconn2odd = dtype_w_skipconn2odd
        self.neurons_per_even_layer = self.sum_edge
        self.neurons_per_odd_layer = self.sum_edge
