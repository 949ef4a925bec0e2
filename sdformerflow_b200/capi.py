"""ctypes binding of libsdf_b200.so (include/sdf_b200.h).

The struct layouts are parsed from the header itself, so the header stays the single source
of truth for the C-ABI.  There is no fallback: if the library is missing, ``lib()`` builds it
with nvcc, and if that fails (or a call returns an error code) a ``RuntimeError`` is raised.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "sdf_b200.h")
LIB_PATH = os.path.join(_HERE, "_lib", "libsdf_b200.so")

# enums of the header, mirrored for python callers
SDF_NEURON_LIF, SDF_NEURON_IF, SDF_NEURON_PLIF = 0, 1, 2
SDF_SPIKE_F32, SDF_SPIKE_U8, SDF_SPIKE_BF16 = 0, 1, 2
SDF_SG_ATAN, SDF_SG_SIGMOID = 0, 1

_SCALARS = {"int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "double": ctypes.c_double,
            "uint8_t": ctypes.c_uint8, "int8_t": ctypes.c_int8, "float": ctypes.c_float}


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse_header(path=HEADER):
    """Returns (structs: {name: [(field, ctype)]}, functions: {name: (restype, [argtype names])})."""
    text = _strip_comments(open(path).read())
    structs, order = {}, []
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl, flags=re.S)
            base, ptr, names = m.group(2), m.group(3), m.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                is_ptr = bool(ptr) or nm.startswith("*")
                nm = nm.lstrip("* ").strip()
                arr = re.match(r"(\w+)\s*\[(\d+)\]$", nm)
                if arr:
                    nm = arr.group(1)
                if is_ptr:
                    ct = ctypes.c_void_p
                elif base in _SCALARS:
                    ct = _SCALARS[base]
                elif base in structs:
                    ct = structs[base]["ctype"]
                else:
                    raise ValueError(f"unknown type {base!r} in struct {name}")
                if arr:
                    ct = ct * int(arr.group(2))
                fields.append((nm, ct))
        cls = type(name, (ctypes.Structure,), {"_fields_": fields})
        structs[name] = {"ctype": cls, "fields": fields}
        order.append(name)
    functions = {}
    for ret, name, args in re.findall(r"\n\s*(int64_t|int|const char\s*\*)\s+(sdf_\w+)\s*\(([^)]*)\)\s*;", text):
        functions[name] = (ret.replace(" ", ""), [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"])
    return structs, functions


_lock = threading.Lock()
_lib = None
_structs = None
_functions = None


def structs():
    global _structs, _functions
    if _structs is None:
        _structs, _functions = parse_header()
    return _structs


def declared_functions():
    structs()
    return _functions


def lib():
    """Loads (building if needed) libsdf_b200.so and sets argtypes/restypes from the header."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # Always go through the incremental, digest-checked build: a stale .so next to an edited header / source would be
        # a silent ABI mismatch (the struct layouts below are parsed from the header at run time).  The build is
        # serialised across processes (torchrun ranks) with a file lock.
        from . import build
        build.build_library_locked()
        L = ctypes.CDLL(LIB_PATH)
        st = structs()
        for name, (ret, args) in declared_functions().items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch: fail loudly
            fn.restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "constchar*": ctypes.c_char_p}[ret]
            argtypes = []
            for a in args:
                m = re.match(r"const\s+(\w+)\s*\*", a)
                if m and m.group(1) in st:
                    argtypes.append(ctypes.POINTER(st[m.group(1)]["ctype"]))
                elif re.match(r"int64_t\s*\*", a):
                    argtypes.append(ctypes.POINTER(ctypes.c_int64))
                elif a.startswith("int64_t"):
                    argtypes.append(ctypes.c_int64)
                else:
                    raise ValueError(f"unsupported argument {a!r} of {name}")
            fn.argtypes = argtypes
        _lib = L
    return _lib


def struct(name, **kw):
    """Instantiate a header struct by name with keyword fields (nested structs accept dicts)."""
    info = structs()[name]
    obj = info["ctype"]()
    ftypes = dict(info["fields"])
    for k, v in kw.items():
        if k not in ftypes:
            raise AttributeError(f"{name} has no field {k}")
        ft = ftypes[k]
        if isinstance(v, dict):
            v = struct(ft.__name__, **v)
        elif ft is ctypes.c_void_p:
            v = None if v is None else ctypes.c_void_p(int(v))
        elif isinstance(v, (list, tuple)):
            v = ft(*v)
        setattr(obj, k, v)
    return obj


class KernelTimer:
    """Optional per-launch device timing (CUDA events on torch's current stream, which is the stream
    every kernel is launched on) + algorithmic bytes, used by bench.py for the roofline line."""

    def __init__(self, only=None):
        self.only = None if only is None else set(only)
        self.records = []

    def wants(self, name):
        return self.only is None or name in self.only

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, nbytes in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
        for d in out.values():
            d["gbps"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        return out


_timer = None


def set_timer(timer):
    global _timer
    _timer = timer


NVTX = os.environ.get("SDF_NVTX", "0") not in ("", "0")


def call(fn_name, args_struct, algo_bytes=0):
    """Calls an ``int sdf_*(const args*)`` entry point; raises with sdf_last_error() on failure."""
    L = lib()
    t = _timer
    if NVTX:
        # SDF_NVTX=1: one NVTX range per C-ABI call, so that a timeline (nsys / torch.profiler) shows the entry points by
        # name around their kernels; off by default (two extra Python calls per launch)
        import torch
        torch.cuda.nvtx.range_push(fn_name)
        try:
            rc = getattr(L, fn_name)(ctypes.byref(args_struct))
        finally:
            torch.cuda.nvtx.range_pop()
    elif t is not None and t.wants(fn_name):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, fn_name)(ctypes.byref(args_struct))
        e1.record()
        t.records.append((fn_name, e0, e1, algo_bytes))
    else:
        rc = getattr(L, fn_name)(ctypes.byref(args_struct))
    if rc != 0:
        raise RuntimeError(f"{fn_name} failed ({rc}): {L.sdf_last_error().decode()}")


def launch_count():
    return int(lib().sdf_launch_count())


def version():
    return int(lib().sdf_version())
