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
    """`gen(word_length, Z, is_y_all_zero=True)` -> (list of LLR arrays [word_length, N*Z], list of codewords), one pair per
    entry of `snr_db`.  The two RandomState streams (noise, information words) are consumed in the reference's order."""

    def __init__(self, N: int, M: int, snr_db: np.ndarray, awgn_noise_seed: int = 2042, wordgen_random_seed: int = 1074,
                 x_dtype=np.float32, y_dtype=np.int64, gen_matrix: np.ndarray = None):
        self.N, self.M, self.K = N, M, N - M
        self.x_dtype, self.y_dtype, self.gen_matrix = x_dtype, y_dtype, gen_matrix
        self.snr_db = snr_db
        self.code_rate = float(self.K) / (N - 2)                         # the reference's rate (Appendix C#3)
        self.snr_lin = np.power(10.0, np.asarray(snr_db) / 10.0)
        self.snr_sigma = np.sqrt(0.5 / (self.snr_lin * self.code_rate))
        self._awgn_noise_random, self._wordgen_random = RandomState(awgn_noise_seed), RandomState(wordgen_random_seed)

    def __call__(self, *args, **kwargs):
        return self._gendata(*args, **kwargs)

    def _gendata(self, word_length: int, Z: int, is_y_all_zero: bool = True):
        if word_length <= 0:
            raise ValueError("word_length must be positive integer")
        n_bits = self.gen_matrix.shape[1]
        llrs, words = [], []
        for sigma in self.snr_sigma:                                     # per SNR point: codewords first, then noise
            words.append(self._codewords(word_length, Z, is_y_all_zero))
            unit = self._awgn_noise_random.normal(0.0, 1.0, size=(word_length, n_bits)).astype(self.x_dtype)
            rx = unit * sigma + -1.0                                     # constant -1: see module docstring
            llrs.append((2 * rx / (sigma ** 2)).astype(self.x_dtype))
        return llrs, words

    def _codewords(self, word_length, Z, all_zero):
        shape = (word_length, self.K * Z)
        if all_zero:
            info = np.zeros(shape, dtype=self.y_dtype)
        elif self.gen_matrix is None:
            raise ValueError("self.gen_matrix must be provided when is_y_all_zero is False")
        else:
            info = self._wordgen_random.randint(0, 2, size=shape).astype(self.y_dtype)
        return np.dot(info, self.gen_matrix) % 2
