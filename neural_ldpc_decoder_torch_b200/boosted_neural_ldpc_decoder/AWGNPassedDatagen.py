"""Boosted AWGNPassedDatagen — BPSK + AWGN LLR generator with the reference's exact random streams and quirks
(/root/reference/src/boosted_neural_ldpc_decoder/AWGNPassedDatagen.py:14-203), vectorised instead of an O(B^2) np.vstack
loop.  Host-side numpy; not on the hot path.  Kept quirks: code rate K/(N - len(punct) -
len(short)) with len(Puncture(0,0)) == 1; `per_snr` fills the whole batch at the first SNR; X is float64."""
import numpy as np
from numpy.random import RandomState

from .Functions import Functions
from .struct.Clipping import Clipping
from .struct.DecoderType import DecoderType
from .struct.Puncture import Puncture
from .struct.Shortening import Shortening


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
            puncturing: Puncture = Puncture(0, 0),
            shortening: Shortening = Shortening(0, 0),
            allowed_llr_range: Clipping = Clipping(abs=20.0),
    ):
        self.N, self.M, self.K = N, M, N - M
        self.snr_db = snr_db
        self.code_rate = 1.0 * self.K / (N - len(puncturing) - len(shortening))
        self.snr_lin = 10.0 ** (self.snr_db / 10.0)
        self.snr_sigma = np.sqrt(1.0 / (2.0 * self.snr_lin * self.code_rate))
        self._awgn_noise_random = RandomState(awgn_noise_seed)
        self._wordgen_random = RandomState(wordgen_random_seed)
        self.x_dtype, self.y_dtype = x_dtype, y_dtype
        self.gen_matrix = gen_matrix
        self.puncturing, self.shortening = puncturing, shortening
        self.allowed_llr_range = allowed_llr_range

    def __call__(self, gentype: str = "per_snr", *args, **kwargs):
        if gentype == "per_snr":
            return self._gendata_per_snr(*args, **kwargs)
        if gentype == "mix_snr":
            return self._gendata_mixed(*args, **kwargs)
        raise AttributeError("attribute `gentype` must be \"per_snr\" or \"mix_snr\".")

    def _generate(self, sigmas, Z, is_y_all_zero, decoding_type, decoder_qms_qbit):
        """one codeword per entry of `sigmas`, drawn in order (same stream as the reference's per-codeword loop)"""
        B, NZ = len(sigmas), self.N * Z
        Y = self._gen_y(B, Z, is_y_all_zero)
        noise = self._awgn_noise_random.normal(0.0, 1.0, (B, NZ))
        sig = np.asarray(sigmas, dtype=np.float64)[:, None]
        received = noise * sig + (-1) ** (1 - Y)           # bit 0 -> -1, bit 1 -> +1
        X = 2 * received / (sig ** 2)
        if decoding_type == DecoderType.QMS:
            X = Functions.Cal_MSA_Q(X, decoder_qms_qbit)
        if self.puncturing.start > 0:
            X[:, self.puncturing.start - 1:self.puncturing.end] = 0.001 if decoding_type == DecoderType.SP else 0
        if self.shortening.start > 0:
            X[:, self.shortening.start - 1:self.shortening.end] = -self.allowed_llr_range.abs   # (reference: AttributeError)
        return np.reshape(X, [B, self.N, Z]), Y

    def _gendata_per_snr(self, word_length: int, Z: int, is_y_all_zero: bool = True,
                         decoding_type: DecoderType = DecoderType.MS, decoder_qms_qbit: int = 5):
        if word_length <= 0:
            raise ValueError("word_length must be positive integer")
        return self._generate([self.snr_sigma[0]] * word_length, Z, is_y_all_zero, decoding_type, decoder_qms_qbit)

    def _gendata_mixed(self, word_length: int, Z: int, is_y_all_zero: bool = True,
                       decoding_type: DecoderType = DecoderType.MS, decoder_qms_qbit: int = 5):
        if word_length <= 0:
            raise ValueError("word_length must be positive integer")
        n = len(self.snr_sigma)
        return self._generate([self.snr_sigma[i % n] for i in range(word_length)], Z, is_y_all_zero, decoding_type,
                              decoder_qms_qbit)

    def _gen_y(self, word_length: int, Z: int, is_y_all_zero: bool) -> np.ndarray:
        if is_y_all_zero:
            return np.zeros([word_length, self.N * Z], dtype=self.y_dtype)
        if self.gen_matrix is None:
            raise ValueError("gen_matrix must be provided when is_y_all_zero is False")
        info = self._wordgen_random.randint(0, 2, size=(word_length, self.K * Z))
        return np.dot(info, self.gen_matrix) % 2
