"""Drop-in mirror of the reference package `boosted_neural_ldpc_decoder`: the package re-exports the data generator, the two
graph classes and `Functions`; the decoder, the loss and the config types are reached through their submodules
(`.BoostedNeuralLDPCDecoder`, `.LDPCDecoderLoss`, `.struct.*`), exactly as the reference's callers import them."""
from . import AWGNPassedDatagen as _datagen, ConnectingMatrix as _cm, ConnectingMatrixTorch as _cmt, Functions as _fn

AWGNPassedDatagen = _datagen.AWGNPassedDatagen
ConnectingMatrix = _cm.ConnectingMatrix
ConnectingMatrixTorch = _cmt.ConnectingMatrixTorch
Functions = _fn.Functions
__all__ = ["AWGNPassedDatagen", "ConnectingMatrix", "ConnectingMatrixTorch", "Functions"]
