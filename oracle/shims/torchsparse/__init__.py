"""TEST INFRASTRUCTURE ONLY — pure-PyTorch CPU restatement of the torchsparse v2.0.0 API surface
that zhen6618/EPRecon's hot path calls (reference call sites: ops/torchsparse_utils.py:5-8,
models/modules.py:3-6, models/neucon_network.py:5, models/gru_fusion.py:3).

torchsparse is an un-vendored third-party dependency of the reference (README.md:17 pins
`-b v2.0.0`); it is not installed or installable in this image, so its *published* algorithm is
restated here.  PARITY UNPINNED: the reference ships no golden vectors for this boundary; every
primitive is cross-checked against dense ATen ops in tests/test_oracle_shims.py instead.

Nothing under eprecon_b200/ may import this package.
"""
from .tensor import PointTensor, SparseTensor  # noqa: F401
from .operators import cat  # noqa: F401
from . import nn  # noqa: F401

__version__ = "2.0.0+oracle-shim"
