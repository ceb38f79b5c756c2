"""spartan2_b200 — B200-native Spartan2 prover hot path.

Host-side mirror (Python, for tests/bench; the product boundary is the C ABI in
include/spartan2_b200.h) of the reference's interface for the path: names and argument meaning
follow microsoft/Spartan2 (`SumcheckProof::prove_cubic_with_three_inputs`, `prove_quad`,
`EqPolynomial::evals_from_points`, `MultilinearPolynomial::bind_poly_var_top`, ...).

Field elements are numpy uint64 arrays of shape (n, 4): little-endian 64-bit limbs in Montgomery form
(R = 2^256) — the reference's in-memory layout (src/big_num/montgomery.rs:17-22)."""
from ._lib import SpartanError, TranscriptState, lib, LIB_PATH  # noqa: F401
from .host import (Comm, CommitmentKey, Context, Keccak256Transcript, NeutronNovaNIFS, PowPolynomial, R1CSWitness, SumcheckRounds,
                   fold_commitments, weights_from_r, DeviceBuffer, DlogGroupExt, EqPolynomial, HyraxPCS, MultilinearPolynomial,
                   SmallValue, SpartanPrepSNARK, SpartanProof, SpartanSNARK, SplitR1CSShape, SumcheckProof, shard_cyclic)  # noqa: F401
