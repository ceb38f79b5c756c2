"""ctypes loader for libspartan2_b200.so (the C ABI declared in include/spartan2_b200.h).

There is no CPU fallback: if the CUDA library is missing, or no sm_100 device is visible, every
entry point raises.  (The CPU restatement under oracle/ is test infrastructure and is never imported
from this package.)"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SP2_LIB_PATH") or os.path.join(HERE, "libspartan2_b200.so")   # (override: A/B measurements of two builds)
_lib = None

OK = 0
ERRORS = {-1: "InternalError(CUDA)", -2: "InvalidInputLength", -3: "InvalidWitnessLength", -4: "InvalidCommitmentKeyLength",
          -5: "DivisionByZero", -6: "InternalError", -7: "Unsupported", -8: "ProofVerifyError"}


class SpartanError(RuntimeError):
    """Mirrors the reference's SpartanError (src/errors.rs:12-110): .kind is the variant name."""

    def __init__(self, code, reason=""):
        self.code = code; self.kind = ERRORS.get(code, "InternalError"); self.reason = reason
        super().__init__("%s: %s" % (self.kind, reason))


class TranscriptState(C.Structure):
    """(round, state[64]) of a Keccak256Transcript right after a squeeze (src/provider/keccak.rs:26-31)."""
    _fields_ = [("round", C.c_uint16), ("state", C.c_uint8 * 64)]

    @classmethod
    def make(cls, state_bytes, rnd):
        t = cls(); t.round = rnd
        for i, b in enumerate(bytes(state_bytes)):
            t.state[i] = b
        return t

    def get(self):
        return bytes(self.state), int(self.round)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("spartan2_b200: %s is missing — run `python -m spartan2_b200.build` (nvcc, sm_100a); "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in header_prototypes().items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = restype, argtypes
    _lib = L
    return L


HEADER_PATH = os.path.join(HERE, "..", "include", "spartan2_b200.h")
_SCALARS = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "uint64_t": C.c_uint64, "int64_t": C.c_int64, "uint16_t": C.c_uint16,
            "uint8_t": C.c_uint8, "int": C.c_int, "float": C.c_float, "double": C.c_double, "size_t": C.c_size_t,
            "sp2_allgather_fn": C.c_void_p}


def header_prototypes(path=HEADER_PATH):
    """{name: (restype, argtypes)} for every function include/spartan2_b200.h declares, so that ctypes converts every
    argument to the width the C side reads (an undeclared uint64_t parameter would travel as a 32-bit int).  Pointers of
    every kind are c_void_p (accepts None, ints, arrays, byref() and callback objects); `const char *` is c_char_p."""
    import re
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)      # struct bodies hold no prototypes
    src = re.sub(r"typedef[^;{]*;", " ", src)
    out = {}

    def ctype(decl, is_ret=False):
        decl = decl.strip()
        if "(" in decl:                                   # function-pointer parameter
            return C.c_void_p
        if "*" in decl:
            base = decl.replace("const", " ").split("*")[0].split()
            if base and base[0] == "char":
                return C.c_char_p
            return C.c_void_p
        toks = [t for t in decl.replace("const", " ").split() if t]
        if toks[0] == "void":
            return None
        return _SCALARS[toks[0]]

    for m in re.finditer(r"([\w\s\*]+?)\b(sp2_\w+)\s*\(([^;{}]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        # split on commas outside parentheses
        parts, depth, cur = [], 0, ""
        for ch in args:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        argtypes = [] if (len(parts) == 1 and parts[0].strip() == "void") else [ctype(a) for a in parts]
        out[name] = (ctype(ret + " ", True), argtypes)
    return out
