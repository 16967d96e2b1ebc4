"""ctypes loader of libdgb200.so (the C-ABI drop-in boundary, include/dgb200.h).

The prototypes are read from the header itself so that the binding cannot drift from the declared ABI.
There is NO fallback: if the shared library is missing the import fails loudly (build it with
`make -C feltor_b200/csrc` or `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGB200_LIB") or os.path.join(_HERE, "libdgb200.so")  # override: kernel experiments only
HEADER_PATH = os.path.join(_HERE, "..", "include", "dgb200.h")

_SCALARS = {
    "int": C.c_int, "double": C.c_double, "size_t": C.c_size_t, "long long": C.c_longlong, "void": None,
    "int32_t": C.c_int32, "int64_t": C.c_int64, "unsigned": C.c_uint,
}


def _ctype(decl):
    """Map a C parameter/return declaration to a ctypes type (every pointer is an opaque address)."""
    decl = decl.strip()
    if decl.endswith("]"):  # array parameter decays to a pointer
        return C.c_void_p
    if "*" in decl or decl.startswith("dgb_stream_t"):
        return C.c_void_p
    toks = [t for t in decl.replace("const", " ").split() if t]
    # drop the parameter name if present
    for n in (2, 1):
        key = " ".join(toks[:n])
        if key in _SCALARS:
            return _SCALARS[key]
    raise ValueError("dgb200.h: cannot map C type of %r" % decl)


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every DGB_API declaration of the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"DGB_API\s+([^;(]+?)\b(dgb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        restype = C.c_char_p if ret.replace(" ", "") == "constchar*" else _ctype(ret + " x" if "*" not in ret else ret)
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        protos[name] = (restype, argtypes)
    return protos


class DgbError(RuntimeError):
    """Raised for a non-zero return code of the C ABI (the C++ shim raises dg::Error / dg::Fail there)."""

    def __init__(self, code, msg):
        super().__init__("dgb error %d: %s" % (code, msg))
        self.code = code


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError("feltor_b200: %s not found - the CUDA library has not been built; there is no "
                              "CPU fallback (run `make -C feltor_b200/csrc`)" % LIB_PATH)
        self.cdll = C.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.raw = {}
        self.missing = []
        for name, (restype, argtypes) in self.protos.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError:  # tests/test_abi.py requires this list to be empty
                self.missing.append(name)
                continue
            fn.restype = restype
            fn.argtypes = argtypes
            self.raw[name] = fn

    def __getattr__(self, name):
        """Checked call: integer-returning entry points raise DgbError on a non-zero code."""
        fn = self.raw["dgb_" + name] if not name.startswith("dgb_") else self.raw[name]
        restype = fn.restype
        if restype is not C.c_int or name in ("version", "dgb_version"):
            return fn

        def call(*args):
            code = fn(*args)
            if code != 0:
                raise DgbError(code, self.raw["dgb_last_error"]().decode())
            return 0
        call.__name__ = name
        setattr(self, name, call)
        return call


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
