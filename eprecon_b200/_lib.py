"""ctypes binding of libeprecon_b200.so — the C-ABI boundary.

include/eprecon_b200.h is the single source of truth: the prototypes are parsed from it, so every declared
symbol must be exported by the library (checked on load) and argument types cannot drift.
No fallback: if the shared library is missing or a call returns non-zero the caller gets an exception.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeprecon_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "eprecon_b200.h")

_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
            "size_t": ctypes.c_size_t, "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32,
            "cudaStream_t": ctypes.c_void_p}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every `int|size_t ep_*(...)` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t)\s+(ep_\w+)\s*\(([^)]*)\)\s*;", src):
        res, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const", "").split()[0]
                    argtypes.append(_SCALARS[ty])
        out[name] = (_SCALARS[res], argtypes)
    return out


_lib = None

# kernels launched per C-ABI call (for bench.py's gpu_launches claim); anything not listed launches one kernel
KERNELS_PER_CALL = {"ep_version": 0, "ep_set_sync_mode": 0, "ep_backproject_workspace_bytes": 0, "ep_compact_workspace_bytes": 0,
                    "ep_sort_segments_workspace_bytes": 0, "ep_spconv_num_row_tiles": 0, "ep_spconv_tc_workspace_bytes": 0, "ep_spconv_hl_workspace_bytes": 0, "ep_hl_slabs": 0, "ep_hl_set_timeline": 0, "ep_spconv_hl_launches": 0, "ep_hl_debug_code": 0, "ep_bn2d_workspace_bytes": 0, "ep_bn2d_train": 2, "ep_backproject_count": 3, "ep_backproject_fused": 3, "ep_backproject_fused_workspace_bytes": 0,
                    "ep_compact_flags": 3, "ep_sort_segments": 14, "ep_hash_build": 2,
                    # native executor calls report their own launch count (executor.py adds it)
                    "ep_masked_attention_workspace_bytes": 0, "ep_masked_attention": 2, "ep_masked_attention_flagged": 2,
                    "ep_exec_decoder_workspace_bytes": 0, "ep_exec_key_tier": 0, "ep_exec_decoder": 44,
                    "ep_exec_launch_count": 0, "ep_exec_profile_enable": 0, "ep_exec_profile_collect": 0, "ep_exec_desc_check": 0, "ep_exec_spvcnn": 0, "ep_exec_gru_level": 0, "ep_exec_linear4x": 0,
                    "ep_exec_init_head": 0}
LAUNCHES = {"n": 0}


class _Counted:
    """ctypes function proxy that counts the kernels each call launches."""

    def __init__(self, fn, k):
        self._fn, self._k = fn, k

    def __call__(self, *a):
        LAUNCHES["n"] += self._k
        return self._fn(*a)



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
        for name, (res, args) in parse_header().items():
            fn = getattr(h, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
            setattr(h, name, _Counted(fn, KERNELS_PER_CALL.get(name, 1)))
        _lib = h
    return _lib


_ERR = {-1: "bad argument", -2: "workspace too small", -3: "CUDA launch error", -4: "unsupported configuration"}


def check(status, what):
    if status != 0:
        raise EpreconError(f"{what} failed: status {status} ({_ERR.get(status, 'unknown')})")
