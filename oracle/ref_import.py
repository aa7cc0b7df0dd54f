"""Import the UNMODIFIED reference hot-path modules from /root/reference on CPU.

Build-container only (the GPU box has no /root/reference).  Three patches make the import possible
without touching the reference tree:
  1. `torch.Tensor.cuda` / `nn.Module.cuda` -> identity (the reference hard-codes `.cuda()`, e.g.
     models/occupancy_initialization.py:74-77, ops/generate_grids.py:8);
  2. `torchsparse` / `spconv.pytorch` resolve to oracle/shims (pure PyTorch restatements);
  3. unrelated, absent imports (skimage, trimesh, pyvista, yacs, ...) resolve to inert stubs.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EPRECON_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_INERT = ("skimage", "trimesh", "matplotlib", "mpl_toolkits", "pyvista", "transforms3d", "yacs",
          "tensorboardX", "memory_profiler", "open3d", "pyrender", "plyfile", "h5py", "ray",
          "numpy_indexed", "torchvision", "cv2", "PIL", "pycuda", "numba", "scipy")


class _Inert(types.ModuleType):
    """Module whose every attribute is another inert object (callable, subscriptable, iterable)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _InertObj(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child


class _InertObj:
    def __init__(self, name="inert"):
        self._n = name

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _InertObj(f"{self._n}.{name}")

    def __call__(self, *a, **k):
        # decorator use: @profile / @numba.njit(...)
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _InertObj):
            return a[0]
        return _InertObj(self._n + "()")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _InertObj(self._n + "[]")


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Inert(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install(force_inert=()):
    """Idempotently install the three patches. Returns the reference root."""
    global _installed
    if _installed:
        return REFERENCE_ROOT
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT} (expected only in the build container)")
    import torch
    from torch import nn

    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self

    missing = []
    for name in _INERT:
        if name in force_inert:
            missing.append(name)
            continue
        try:
            importlib.import_module(name)
        except Exception:
            missing.append(name)
    sys.meta_path.append(_InertFinder(missing))
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    _installed = True
    return REFERENCE_ROOT


def load():
    """Return a namespace with the reference symbols on the hot path."""
    install()
    ns = types.SimpleNamespace()
    ns.back_project = importlib.import_module("ops.back_project").back_project
    ns.generate_grid = importlib.import_module("ops.generate_grids").generate_grid
    ns.tsutils = importlib.import_module("ops.torchsparse_utils")
    ns.modules = importlib.import_module("models.modules")
    ns.occ = importlib.import_module("models.occupancy_initialization")
    ns.gru = importlib.import_module("models.gru_fusion")
    ns.neucon = importlib.import_module("models.neucon_network")
    return ns
