"""CPU oracle for the neural-BP LDPC decode path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (neural_ldpc_decoder_torch_b200) never does.
"""
from .oracle import (build, lib, neural_forward, neural_forward_last, boosted_forward, boosted_step, pack_hard, quantize, count_errors,
                     count_errors_packed)  # noqa: F401
