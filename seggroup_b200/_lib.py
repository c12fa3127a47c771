"""ctypes binding of libseggroup_b200.so — the only way the Python layer reaches the CUDA kernels.

Prototypes are parsed from `include/seggroup_b200.h` (single source of truth for the C-ABI), so a
symbol declared there but missing from the library fails at load time, loudly.  There is NO fallback:
if the library is absent or a kernel call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "seggroup_b200.h")
LIB_PATH = os.environ.get("SGB_LIB_PATH") or os.path.join(_HERE, "lib", "libseggroup_b200.so")     # override: kernel experiments (tools/ablate.sh)

_SCALARS = {"unsigned long long": ctypes.c_ulonglong, "unsigned": ctypes.c_uint, "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
            "long long": ctypes.c_longlong}


class SgbError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> dict:
    """{name: (restype, [(ctype, argname), ...])} for every `sgb_*` function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"\b(long long|int|size_t|const char\s*\*)\s+(sgb_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "long long": ctypes.c_longlong}.get(ret, ctypes.c_char_p)
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argl.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    ty, nm = a.rsplit(" ", 1)
                    ty = ty.replace("const ", "").strip()
                    argl.append((_SCALARS[ty], nm))
        protos[name] = (restype, argl)
    return protos


_lib = None
_protos = None


def load():
    """Load the library (building it with nvcc first if the sources are newer and nvcc exists)."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH) or os.environ.get("SGB_REBUILD"):
        from . import build as _build
        _build.build()
    if not os.path.isfile(LIB_PATH):
        raise SgbError("libseggroup_b200.so not found at %s (run `python -m seggroup_b200.build`)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argl) in _protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SgbError("symbol %s declared in %s is missing from %s" % (name, HEADER, LIB_PATH)) from e
        fn.restype = restype
        fn.argtypes = [t for t, _ in argl]
    _lib = lib
    return lib


def prototypes() -> dict:
    load()
    return _protos


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if isinstance(x, bytes):
        return ctypes.cast(ctypes.c_char_p(x), ctypes.c_void_p)
    if hasattr(x, "ctypes"):                       # host numpy array (only for *_host entry points)
        return ctypes.c_void_p(x.ctypes.data)
    return ctypes.c_void_p(x.data_ptr())          # torch tensor


# optional per-entry-point device timing (tools/profile_scene.py): name -> [calls, ms]
profile = None

# bench.py bookkeeping: CUDA-event pairs around one chosen entry point (recorded on the launching stream, resolved by the
# caller after its final synchronize).  Kernel launches are counted inside the library (sgb_launch_count).
time_entry = None          # set of entry-point names whose calls are bracketed by CUDA events
timed_events = []


def launch_count() -> int:
    return int(load().sgb_launch_count())


def enable_profile():
    global profile
    profile = {}


_ws_cache = {}     # (name, sizes) -> bytes: the *_ws_bytes entry points are pure functions of their integer arguments


def call(name: str, *args):
    if name.endswith("_ws_bytes"):
        key = (name,) + args
        v = _ws_cache.get(key)
        if v is None:
            v = _ws_cache[key] = _call(name, *args)
        return v
    if time_entry is not None and name in time_entry:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _call(name, *args)
        e1.record()
        # tag: two_layer flag for the EdgeConv forward / backward, number of feature rows for the pooling
        tag = args[3] if name == "sgb_edgeconv_fwd" else (args[7] if name == "sgb_edgeconv_bwd" else
                                                          (args[1] if name == "sgb_segment_pool_max_fwd" else 0))
        timed_events.append((name, tag, e0, e1))
        return rc
    if profile is not None and not name.endswith("_bytes"):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _call(name, *args)
        e1.record()
        e1.synchronize()
        rec = profile.setdefault(name, [0, 0.0])
        rec[0] += 1
        rec[1] += e0.elapsed_time(e1)
        return rc
    return _call(name, *args)


_bound = {}        # name -> (function pointer, pointer-argument mask, argument count, checks the status?)


def _bind(name):
    lib = load()
    restype, argl = _protos[name]
    fn = getattr(lib, name)
    b = (fn, tuple(t is ctypes.c_void_p for t, _ in argl), len(argl),
         restype is ctypes.c_int and name not in ("sgb_version", "sgb_last_cuda_error"))
    _bound[name] = b
    return b


def _call(name: str, *args):
    """Call `name`; tensors / None / ints are passed as pointers where the prototype has a pointer.
    Raises SgbError on a negative status.  Returns the int / size_t result."""
    b = _bound.get(name)
    if b is None:
        b = _bind(name)
    fn, is_ptr, n, checked = b
    if len(args) != n:
        raise TypeError("%s expects %d arguments (%s), got %d" % (name, n, ", ".join(nm for _, nm in _protos[name][1]), len(args)))
    conv = [None] * n
    for i in range(n):
        a = args[i]
        if is_ptr[i] and a is not None and a.__class__ is not int:
            a = a.data_ptr() if hasattr(a, "data_ptr") else _ptr(a)       # torch tensor -> raw device address (a plain int)
        conv[i] = a
    rc = fn(*conv)
    if checked and rc < 0:
        lib = load()
        msg = lib.sgb_status_string(rc).decode()
        if rc == -3:
            msg += ": " + lib.sgb_last_cuda_error_string().decode()
        raise SgbError("%s failed: %s (%d)" % (name, msg, rc))
    return rc
