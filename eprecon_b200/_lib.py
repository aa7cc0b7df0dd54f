"""ctypes binding of libeprecon_b200.so — the C-ABI boundary (include/eprecon_b200.h).

No fallback: if the shared library is missing or a call returns a non-zero status the caller gets
an exception.  Signatures carry only raw pointers / sizes / a cudaStream_t.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeprecon_b200.so")

_P, _I, _I64, _F, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/eprecon_b200.h one to one
SIGNATURES = {
    "ep_version": (_I, []),
    "ep_backproject_workspace_bytes": (_SZ, [_I64]),
    "ep_backproject_count": (_I, [_P, _I64, _P, _F, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "ep_backproject_compact": (_I, [_P, _P, _I64, _I, _P, _P, _P, _P, _P]),
    "ep_backproject_gather": (_I, [_P, _P, _I64, _P, _I, _I, _I, _I, _I, _P, _F, _P, _I, _P, _I, _P, _P]),
    "ep_backproject_grid": (_I, [_P, _P, _I64, _I, _I, _I, _I, _P, _F, _P, _P, _P, _P]),
    "ep_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _P]),
}

_lib = None


class EpreconError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EpreconError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the eprecon_b200 hot path)")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


_ERR = {-1: "bad argument", -2: "workspace too small", -3: "CUDA launch error", -4: "unsupported configuration"}


def check(status, what):
    if status != 0:
        raise EpreconError(f"{what} failed: status {status} ({_ERR.get(status, 'unknown')})")
