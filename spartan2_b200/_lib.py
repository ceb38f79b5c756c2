"""ctypes loader for libspartan2_b200.so (the C ABI declared in include/spartan2_b200.h).

There is no CPU fallback: if the CUDA library is missing, or no sm_100 device is visible, every
entry point raises.  (The CPU restatement under oracle/ is test infrastructure and is never imported
from this package.)"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libspartan2_b200.so")
_lib = None

OK = 0
ERRORS = {-1: "InternalError(CUDA)", -2: "InvalidInputLength", -3: "InvalidWitnessLength", -4: "InvalidCommitmentKeyLength",
          -5: "DivisionByZero", -6: "InternalError", -7: "Unsupported"}


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
    L.sp2_last_error.restype = C.c_char_p
    L.sp2_launch_count.restype = C.c_uint64
    L.sp2_sc_tail_len.restype = C.c_uint64
    L.sp2_neutronnova_prep_free.restype = None
    L.sp2_neutronnova_prep_free.argtypes = [C.c_void_p]
    _lib = L
    return L
