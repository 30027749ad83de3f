"""B200-native (sm_100a) neural belief-propagation LDPC decode: a drop-in for the hot path of
ShapeLayer/neural-ldpc-decoder-torch (NeuralLDPCDecoder / BoostedNeuralLDPCDecoder forward + backward).

Sub-packages mirror the reference's import surface:
    neural_ldpc_decoder_torch_b200.neural_ldpc_decoder          (reference: src/neural_ldpc_decoder)
    neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder  (reference: src/boosted_neural_ldpc_decoder)
    neural_ldpc_decoder_torch_b200.checkpoint_utils             (reference: src/checkpoint_utils)
`install_dropin()` additionally registers them under the reference's top-level names.
"""
import importlib
import sys

from .graph import TannerGraph, load_basegraph  # noqa: F401

__all__ = ["TannerGraph", "load_basegraph", "install_dropin"]

_DROPIN = ("neural_ldpc_decoder", "boosted_neural_ldpc_decoder", "checkpoint_utils")


def install_dropin():
    """Make `import neural_ldpc_decoder`, `import boosted_neural_ldpc_decoder`, `import checkpoint_utils`
    resolve to this package's implementations (same class names / signatures as the reference)."""
    for name in _DROPIN:
        mod = importlib.import_module(f"{__name__}.{name}")
        sys.modules[name] = mod
        prefix = f"{__name__}.{name}."
        for k, v in list(sys.modules.items()):
            if k.startswith(prefix):
                sys.modules[name + "." + k[len(prefix):]] = v
