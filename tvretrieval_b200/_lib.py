"""ctypes binding of libxmlb200.so (the C ABI declared in include/xmlb200.h).

The prototypes are parsed from the header so that the binding can never drift from the declared ABI.
There is no fallback: if the shared library is missing, `lib()` raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxmlb200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "xmlb200.h")

_CTYPE = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float,
          "double": ctypes.c_double}
_lib = None


def parse_header(path=HEADER):
    """-> {function name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"(const char\*|int|long long)\s+(xmlb_\w+)\s*\(([^)]*)\)\s*;", text):
        argtypes = []
        for a in [a.strip() for a in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                base = a.rsplit(" ", 1)[0].replace("const ", "").replace("unsigned ", "").strip()
                argtypes.append(_CTYPE[base])
        restype = {"const char*": ctypes.c_char_p, "int": ctypes.c_int, "long long": ctypes.c_longlong}[ret]
        protos[name] = (restype, argtypes)
    return protos


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libxmlb200.so is not built (%s). Run `python -m tvretrieval_b200.build`; this package has no "
                "CPU or PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


class XmlbError(RuntimeError):
    pass


def check(rc, name):
    if rc == 0:
        return
    msg = lib().xmlb_last_error().decode()
    if rc < 0:
        # invalid argument / unsupported shape: same exception family torch raises for bad shapes
        raise XmlbError("%s: %s" % (name, msg))
    raise XmlbError("%s: CUDA error %d: %s" % (name, rc, msg))


def launch_count():
    return int(lib().xmlb_launch_count())
