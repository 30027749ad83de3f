"""AWGNPassedDatagen — BPSK + AWGN channel LLR generator with the reference's exact random stream
(/root/reference/src/neural_ldpc_decoder/AWGNPassedDatagen.py:5-98): same seeds -> same tensors, so
"identical seeded inputs" means the same thing for both implementations.  Host-side numpy; not on the hot path.

Reference quirks kept on purpose: the modulation term is the constant -1
(the reference's `-1 ** (1 - y)` parses as -(1 ** ...)), i.e. only correct for the all-zero codeword,
and the code rate is K / (N - 2).
"""
import numpy as np
from numpy.random import RandomState


class AWGNPassedDatagen:
    def __init__(
            self,
            N: int,
            M: int,
            snr_db: np.ndarray,
            awgn_noise_seed: int = 2042,
            wordgen_random_seed: int = 1074,
            x_dtype=np.float32,
            y_dtype=np.int64,
            gen_matrix: np.ndarray = None,
    ):
        self.N, self.M, self.K = N, M, N - M
        self.snr_db = snr_db
        self.code_rate = 1.0 * (N - M) / (N - 2)
        self.snr_lin = 10.0 ** (self.snr_db / 10.0)
        self.snr_sigma = np.sqrt(1.0 / (2.0 * self.snr_lin * self.code_rate))
        self._awgn_noise_random = RandomState(awgn_noise_seed)
        self._wordgen_random = RandomState(wordgen_random_seed)
        self.x_dtype, self.y_dtype = x_dtype, y_dtype
        self.gen_matrix = gen_matrix

    def __call__(self, *args, **kwargs):
        return self._gendata(*args, **kwargs)

    def _gendata(self, word_length: int, Z: int, is_y_all_zero: bool = True):
        if word_length <= 0:
            raise ValueError("word_length must be positive integer")
        xs, ys = [], []
        for sigma in self.snr_sigma:
            y = self._codewords(word_length, Z, is_y_all_zero)
            noise = self._awgn_noise_random.normal(0., 1., size=(word_length, self.gen_matrix.shape[1])).astype(self.x_dtype)
            received = noise * sigma + -1.0     # constant -1: see module docstring
            xs.append((2 * received / (sigma ** 2)).astype(self.x_dtype))
            ys.append(y)
        return xs, ys

    def _codewords(self, word_length, Z, all_zero):
        if all_zero:
            info = np.zeros(shape=(word_length, self.K * Z), dtype=self.y_dtype)
        else:
            if self.gen_matrix is None:
                raise ValueError("self.gen_matrix must be provided when is_y_all_zero is False")
            info = self._wordgen_random.randint(0, 2, size=(word_length, self.K * Z)).astype(self.y_dtype)
        return np.dot(info, self.gen_matrix) % 2
