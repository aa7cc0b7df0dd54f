"""eprecon_b200 — B200-native (sm_100a) implementation of the EPRecon feature-volume hot path.

Host side mirrors the reference's nn.Module interface (same names, forward() signatures, state-dict
layout); every hot op calls hand-written CUDA in `libeprecon_b200.so` through a C ABI
(include/eprecon_b200.h).  There is no CPU fallback: ops raise if the library is missing.
"""
__version__ = "0.1.0"
