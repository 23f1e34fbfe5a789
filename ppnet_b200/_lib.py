"""ctypes loader for the C-ABI library (ppnet_b200/lib/libppnet_b200.so, built by
`python -c "import __graft_entry__ as g; g.build()"` or `make -C ppnet_b200/csrc`).

There is NO CPU fallback: if the library is missing or a call fails, an exception is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPNET_B200_LIB: development override (A/B builds of the same library); the product always loads the in-tree file
LIB_PATH = os.environ.get("PPNET_B200_LIB") or os.path.join(_HERE, "lib", "libppnet_b200.so")
_lib = None


class PPNetError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PPNetError(
                "ppnet_b200: CUDA library %s is missing -- build it with `make -C ppnet_b200/csrc` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ppnet_last_error.restype = ctypes.c_char_p
        _lib.ppnet_launch_count.restype = ctypes.c_int64
        _lib.ppnet_version.restype = ctypes.c_int
    return _lib


def check(rc, what):
    if rc != 0:
        raise PPNetError("%s failed (%d): %s" % (what, rc, lib().ppnet_last_error().decode()))


def launch_count():
    return int(lib().ppnet_launch_count())


def exported_symbols():
    """Names declared in include/ppnet_b200.h (parsed) -- used by the CPU test that the library
    exports everything the header promises."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "ppnet_b200.h")
    with open(hdr) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ppnet_[a-z0-9_]+)\s*\(", txt)))
