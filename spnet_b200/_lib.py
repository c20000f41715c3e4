"""ctypes binding of libspnet_b200.so (the C ABI declared in include/spnet_b200.h).

The signatures are parsed from the header itself, so the Python side cannot drift from
the C side. There is no fallback: if the shared library is missing or a call fails, this
raises. Pointers are passed as integers (torch ``tensor.data_ptr()``) or None.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libspnet_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "spnet_b200.h")


class SpnetError(RuntimeError):
    pass


_SCALARS = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "cudaStream_t": ctypes.c_void_p,
}


def _ctype_of(decl):
    decl = decl.replace("const ", "").strip()
    if "*" in decl:
        return ctypes.c_void_p
    # drop the parameter name
    parts = decl.split()
    tname = " ".join(parts[:-1]) if len(parts) > 1 else parts[0]
    if tname.startswith("unsigned"):
        raise ValueError("unsupported by-value type: " + decl)
    return _SCALARS[tname]


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"^(int|const char\*)\s+(spnet_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (ctypes.c_int if ret == "int" else ctypes.c_char_p, argtypes)
    return protos


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise SpnetError(
                "libspnet_b200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self._dll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        self.launches = 0  # kernels launched through this binding (bench.py reports it)
        self.profile = None  # list -> (name, args, start_event, end_event) per launch (bench.py)

    def last_error(self):
        return self._dll.spnet_last_error().decode()

    def __getattr__(self, name):
        # spnet_<name> with return-code checking
        full = name if name.startswith("spnet_") else "spnet_" + name
        fn = getattr(self._dll, full)
        if full in ("spnet_version", "spnet_last_error"):
            return fn

        def call(*args):
            if self.profile is not None:
                import torch
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args)
                e1.record()
                self.profile.append((full, args, e0, e1))
            else:
                rc = fn(*args)
            if rc != 0:
                raise SpnetError("%s failed (%d): %s" % (full, rc, self.last_error()))
            self.launches += 1
            return rc

        call.__name__ = full
        setattr(self, name, call)
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
        import torch
        if torch.cuda.is_available():  # refuse anything but sm_100a up front (the library carries no other code)
            rc = _lib._dll.spnet_check_device()
            if rc != 0:
                raise SpnetError("spnet_check_device failed (%d): %s" % (rc, _lib.last_error()))
    return _lib
