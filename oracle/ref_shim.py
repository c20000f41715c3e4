"""ORACLE tooling — authoring-container only (never runs on the GPU box, never imported by the
product). Makes the reference's numpy-side modules importable from /root/reference without
Keras / TensorFlow / matplotlib, following SURVEY.md appendix A: a meta-path finder that
fabricates permissive stub modules for those packages, plus numpy.int for numpy >= 1.24."""
import importlib.abc
import importlib.machinery
import sys
import types

REFERENCE_ROOT = "/root/reference"
STUBBED = ("keras", "tensorflow", "matplotlib", "numba")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        # used as a decorator (@jit) -> hand the function back; otherwise another stub
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], (_Anything, type)):
            return a[0]
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (_StubBase,)

    def __iter__(self):
        return iter(())


class _StubBase:
    def __init__(self, *a, **k):
        pass


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        v = _Anything()
        setattr(self, name, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
