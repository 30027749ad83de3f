"""Drop-in mirror of the reference package `neural_ldpc_decoder` (same four public names; the decoder runs on the sm_100a
kernels behind include/nldpc.h).  `import neural_ldpc_decoder.NeuralLDPCDecoder as X` yields the CLASS, as with the reference,
because the re-exported names shadow the submodules."""
from . import AWGNPassedDatagen as _datagen, ConnectingMatrix as _cm, ConnectingMatrixTorch as _cmt, NeuralLDPCDecoder as _dec

AWGNPassedDatagen = _datagen.AWGNPassedDatagen
ConnectingMatrix = _cm.ConnectingMatrix
ConnectingMatrixTorch = _cmt.ConnectingMatrixTorch
NeuralLDPCDecoder = _dec.NeuralLDPCDecoder
__all__ = ["AWGNPassedDatagen", "ConnectingMatrix", "ConnectingMatrixTorch", "NeuralLDPCDecoder"]
