"""ctypes binding of ``libganffn.so`` (the C ABI declared in ``include/ganffn.h``).

The prototypes are read from the header itself, so the binding cannot drift from
the declarations.  There is no CPU fallback: if the shared library is missing the
import of any compute entry point raises, loudly (``__graft_entry__.build()`` or
``make -C gan_ffn_b200/csrc`` produces it).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ganffn.h")
# GANFFN_LIB: tuning aid only (A/B runs of a variant build made by tools/build_variant.sh)
LIB_PATH = os.environ.get("GANFFN_LIB") or os.path.join(_HERE, "libganffn.so")

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "unsigned long long": ctypes.c_ulonglong,
    "double": ctypes.c_double,
    "void": None,
}

_PROTO = re.compile(r"^(const char\*|unsigned long long|int64_t|int|void)\s+(ganffn_\w+)\s*\(([^;]*?)\)\s*;", re.M | re.S)


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """name -> (restype, argtypes) for every function declared in the header."""
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    out: Dict[str, Tuple[object, List[object]]] = {}
    for ret, name, args in _PROTO.findall(text):
        restype = ctypes.c_char_p if ret == "const char*" else _CTYPES[ret]
        argtypes: List[object] = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
        out[name] = (restype, argtypes)
    return out


class GanffnError(RuntimeError):
    pass


class _Lib:
    def __init__(self) -> None:
        if not os.path.exists(LIB_PATH):
            raise GanffnError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C gan_ffn_b200/csrc`. There is no CPU fallback for this path.")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes

    def last_error(self) -> str:
        return self.cdll.ganffn_last_error().decode()

    def call(self, name: str, *args):
        """Calls a status-returning entry point and raises on a non-zero status."""
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            kind = ValueError if rc == 1 else GanffnError
            raise kind(f"{name} failed (status {rc}): {self.last_error()}")

    def query(self, name: str, *args) -> int:
        """Calls a size query; -1 means a precondition failed."""
        v = getattr(self.cdll, name)(*args)
        if v < 0:
            raise ValueError(f"{name}: {self.last_error()}")
        return int(v)


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def ptr(t) -> int:
    """Device pointer of a tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()
