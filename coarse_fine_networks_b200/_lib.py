"""ctypes binding of libcfnet_b200.so (the C ABI declared in include/cfnet_b200.h).

There is no fallback: if the shared library is missing, importing this module raises, and
every op in the package goes through it."""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CFNET_LIB") or os.path.join(_PKG, "libcfnet_b200.so")     # CFNET_LIB: another build of the same library (A/B runs)
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "cfnet_b200.h")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -m coarse_fine_networks_b200.build` "
        "(nvcc, sm_100a). coarse_fine_networks_b200 has no CPU / PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)
lib.cf_last_error.restype = ctypes.c_char_p
lib.cf_launch_count.restype = ctypes.c_ulonglong
lib.cf_abi_version.restype = ctypes.c_int

_c_int, _c_i64, _c_f32, _c_ptr, _c_size = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_CTYPES = {"int": _c_int, "int64_t": _c_i64, "float": _c_f32, "size_t": _c_size, "cudaStream_t": _c_ptr,
           "double": ctypes.c_double, "unsigned long long": ctypes.c_ulonglong}


def _parse_header():
    """Read the prototypes out of include/cfnet_b200.h so that the binding cannot drift from the ABI."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|size_t|const char\*|unsigned long long)\s+(cf_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(_c_ptr)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPES[ty])
        protos[name] = (ret, argtypes)
    return protos


PROTOS = _parse_header()
for _name, (_ret, _argtypes) in PROTOS.items():
    _fn = getattr(lib, _name)            # AttributeError here == header/library mismatch
    _fn.argtypes = _argtypes
    _fn.restype = {"int": _c_int, "size_t": _c_size, "const char*": ctypes.c_char_p,
                   "unsigned long long": ctypes.c_ulonglong}[_ret]


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a dense CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("cfnet_b200: tensor is not on a CUDA device (no CPU fallback)")
    return t.data_ptr()


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError(cf_last_error()) on failure."""
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.cf_last_error().decode()}")


def launch_count():
    return int(lib.cf_launch_count())


# ----------------------------------------------------------------------------------------
# ctypes mirrors of the argument structs, generated from the header so they cannot drift
# ----------------------------------------------------------------------------------------
def _parse_structs():
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef struct\s*\{(.*?)\}\s*(cf_\w+)\s*;", txt, flags=re.S):
        body, name = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            is_ptr = "*" in decl
            decl = decl.replace("*", " ")
            toks = decl.replace(",", " ").split()
            names = []
            while toks and (toks[-1] not in _CTYPES and toks[-1] not in structs and toks[-1] != "const"):
                names.insert(0, toks.pop())
            ty = " ".join(t for t in toks if t != "const")
            if is_ptr:
                ct = _c_ptr
            elif ty in structs:
                ct = structs[ty]
            else:
                ct = _CTYPES[ty]
            fields += [(n, ct) for n in names]
        structs[name] = type(name, (ctypes.Structure,), {"_fields_": fields})
    return structs


STRUCTS = _parse_structs()
for _sname in STRUCTS:
    if ("cf_sizeof_" + _sname[3:]) not in PROTOS:
        continue
    _sz = getattr(lib, "cf_sizeof_" + _sname[3:])()
    if _sz != ctypes.sizeof(STRUCTS[_sname]):
        raise ImportError(f"ABI mismatch for {_sname}: library {_sz} bytes, header {ctypes.sizeof(STRUCTS[_sname])}")


def make(struct_name, **kw):
    """Build an argument struct; tensors are converted to device pointers, None to NULL."""
    s = STRUCTS[struct_name]()
    for k, v in kw.items():
        if torch.is_tensor(v):
            v = ptr(v)
        setattr(s, k, v)
    return s


def call_struct(name, s):
    rc = getattr(lib, name)(ctypes.byref(s), stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.cf_last_error().decode()}")
