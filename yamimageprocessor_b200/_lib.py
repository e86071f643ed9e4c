"""ctypes binding of ``libyamb200.so`` (C ABI declared in ``include/yamb200.h``).

The library is the product: there is no CPU fallback.  ``load()`` raises
``BackendUnavailable`` with build instructions when the shared object is
missing, and ``Context`` raises ``YamError`` carrying the library's
``yam_last_error()`` text for every non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading
from pathlib import Path
from typing import Dict, List, Optional, Tuple

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
HEADER = REPO_ROOT / "include" / "yamb200.h"
LIB_PATH = PKG_DIR / "libyamb200.so"

YAM_U8, YAM_U16, YAM_F32, YAM_I32 = 0, 1, 2, 3
BORDER_REFLECT101, BORDER_REPLICATE = 0, 1
MORPH_ERODE, MORPH_DILATE, MORPH_OPEN, MORPH_CLOSE = 0, 1, 2, 3
SHAPE_RECT, SHAPE_ELLIPSE, SHAPE_CROSS = 0, 1, 2
PROPS_STRIDE = 8


class BackendUnavailable(RuntimeError):
    """libyamb200.so is not built / cannot be loaded."""


class YamError(RuntimeError):
    """A libyamb200 call returned an error code."""

    def __init__(self, func: str, code: int, message: str) -> None:
        super().__init__(f"{func} failed ({code}): {message}")
        self.func = func
        self.code = code
        self.message = message


_CTYPES = {
    "int": C.c_int,
    "int64_t": C.c_int64,
    "double": C.c_double,
    "float": C.c_float,
    "void": None,
}


def _ctype_of(decl: str):
    """Map a C parameter/return declaration from the header to a ctypes type."""
    d = decl.replace("const", " ").strip()
    stars = d.count("*")
    base = d.replace("*", " ").split()
    # drop the parameter name when present
    known = {"int", "int64_t", "int32_t", "uint8_t", "uint16_t", "uint32_t", "uint64_t", "float", "double", "void", "char",
             "yam_ctx"}
    tokens = [t for t in base if t in known]
    if not tokens:
        raise ValueError(f"cannot parse C declaration: {decl!r}")
    t = tokens[0]
    if stars == 0:
        return _CTYPES[t]
    if t == "char" and stars == 1:
        return C.c_char_p
    return C.c_void_p  # every pointer crosses the boundary as an opaque address


def parse_header(path: Path = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """Return {symbol: (restype, argtypes)} for every function the header declares."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    protos: Dict[str, Tuple[object, List[object]]] = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(yam_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        restype = _ctype_of(ret)
        argtypes: List[object] = []
        if args and args != "void":
            for a in args.split(","):
                argtypes.append(_ctype_of(a))
        protos[name] = (restype, argtypes)
    return protos


_lock = threading.Lock()
_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libyamb200.so and attach prototypes parsed from the header."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = Path(os.environ.get("YAM_B200_LIB", str(LIB_PATH)))
        if not path.exists():
            raise BackendUnavailable(
                f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {PKG_DIR / 'csrc'}`; this backend has no CPU fallback."
            )
        try:
            lib = C.CDLL(str(path))
        except OSError as exc:  # missing libcudart etc.
            raise BackendUnavailable(f"cannot load {path}: {exc}") from exc
        for name, (restype, argtypes) in parse_header().items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise BackendUnavailable(f"{path} does not export {name} (stale build?)") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        abi = lib.yam_abi_version()
        if abi != 1:
            raise BackendUnavailable(f"{path} has ABI version {abi}, expected 1")
        _lib = lib
        return lib


def last_error() -> str:
    msg = load().yam_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(func: str, code: int) -> None:
    if code != 0:
        raise YamError(func, code, last_error())
