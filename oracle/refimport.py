"""Import the reference's own modules (read-only at /root/reference) unchanged, in THIS container only.

The reference imports a dozen third-party packages that are not installed.  A
sys.meta_path finder fabricates inert stub modules for those roots, and the six symbols the hot path actually
executes are replaced by the restatements in oracle/third_party.py.  Used by oracle/make_golden.py (fixture
generation) and by the CPU tests that cross-check the oracle restatements against the live reference; the GPU box
has no /root/reference, so nothing that runs there may call `load()`.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("REHRSEG_REFERENCE", "/root/reference")

_STUB_ROOTS = ("dynamic_network_architectures", "nnunetv2", "acvl_utils", "batchgenerators", "batchgeneratorsv2",
               "resize", "degrade", "kornia", "nibabel", "SimpleITK", "h5py", "omegaconf", "sigpy", "skimage")


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(f"{cls.__name__}.{name}")


def _make_stub(name: str):
    def _init(self, *a, **k):
        pass

    def _call(self, *a, **k):
        raise RuntimeError(f"oracle stub `{name}` was executed: this third-party symbol needs a real restatement")

    return _StubMeta(name.replace(".", "_"), (object,), {"__init__": _init, "__call__": _call, "_stub_name": name})


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        stub = _make_stub(f"{self.__name__}.{name}")
        setattr(self, name, stub)
        return stub


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install() -> None:
    """Install the stub finder + real shims and put the reference root on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (it exists only in the build container)")
    real = [r for r in _STUB_ROOTS if importlib.util.find_spec(r) is not None]
    # anything genuinely installed wins over a stub
    finder = _StubFinder()
    finder_roots = tuple(r for r in _STUB_ROOTS if r not in real)
    finder.find_spec = (lambda fullname, path=None, target=None, _f=finder, _r=finder_roots:
                        importlib.machinery.ModuleSpec(fullname, _f, is_package=True)
                        if fullname.split(".")[0] in _r else None)
    sys.meta_path.append(finder)
    from . import third_party as tp

    def put(modname, **symbols):
        mod = importlib.import_module(modname)
        for k, v in symbols.items():
            setattr(mod, k, v)

    put("dynamic_network_architectures.architectures.unet", PlainConvUNet=tp.PlainConvUNet)
    put("dynamic_network_architectures.building_blocks.unet_decoder", UNetDecoder=tp.UNetDecoder)
    put("nnunetv2.inference.sliding_window_prediction", compute_gaussian=tp.compute_gaussian)
    put("acvl_utils.cropping_and_padding.padding", pad_nd_image=tp.pad_nd_image)
    put("nnunetv2.training.loss.dice", SoftDiceLoss=tp.SoftDiceLoss, MemoryEfficientSoftDiceLoss=tp.MemoryEfficientSoftDiceLoss)
    put("nnunetv2.utilities.helpers", softmax_helper_dim1=tp.softmax_helper_dim1)
    put("nnunetv2.training.loss.deep_supervision", DeepSupervisionWrapper=tp.DeepSupervisionWrapper)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load(name: str):
    """importlib.import_module of a reference module, e.g. load('models.seg_model') or load('utils.fba')."""
    install()
    return importlib.import_module(name)
