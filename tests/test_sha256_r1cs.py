"""The SHA-256 -> R1CS workload generator (spartan2_b200/frontend) — CPU only.
Checkpoints from the reference: 25,840 constraints per compression beyond the 512 input-bit constraints
(benches/sha256_neutronnova.rs:159-160: 26,352 in total), padded shapes per SplitR1CSShape::new
(src/r1cs/mod.rs:810-911), digest equal to SHA-256 of the message (benches/sha256_spartan.rs:104-118)."""
import hashlib

import numpy as np
import pytest

from spartan2_b200.frontend import Sha256Circuit


def test_compression_constraint_count_matches_reference_quote():
    c = Sha256Circuit(bytes(range(64)), kind="compression")
    assert c.num_cons_unpadded - 512 == 25840
    assert c.num_cons_unpadded == 26352 and c.num_cons == 32768          # "~26,352 constraints ... padded to 32,768"
    assert c.is_satisfied()


@pytest.mark.parametrize("n", [0, 1, 55, 56, 64, 100])
def test_digest_and_satisfiability(n):
    msg = bytes((7 * i + 3) % 256 for i in range(n))
    c = Sha256Circuit(msg)
    assert c.digest == hashlib.sha256(msg).digest()
    assert c.is_satisfied()
    assert c.num_public == 256
    # public inputs are the digest bits, big-endian per byte (sha256_spartan.rs:56-73)
    assert bytes(np.packbits(c.pub_bits)) == c.digest
    # a flipped witness bit breaks it
    if len(c.aux_bits):        # (the empty message constant-folds to no witness at all)
        c.aux_bits[len(c.aux_bits) // 2] ^= 1
        assert not c.is_satisfied()


def test_padded_shape_rules():
    c = Sha256Circuit(b"\x00" * 64)
    assert c.num_precommitted % 2048 == 0 and c.num_precommitted >= c.num_aux
    nv = c.num_precommitted + c.num_rest
    assert nv & (nv - 1) == 0 and c.num_cons & (c.num_cons - 1) == 0
    for (coef, idx, ptr) in c.raw:
        assert ptr.shape[0] == c.num_cons + 1 and int(idx.max()) < nv + 1 + 256
        # columns of padded witness slots are never referenced
        assert not np.any((idx >= c.num_aux) & (idx < nv))


def test_bench_config_sizes_1kib():
    # BASELINE config 1: 1 KiB of zeros -> 17 blocks; SURVEY §8 estimate N = M = 2^19
    c = Sha256Circuit(b"\x00" * 1024)
    assert c.digest == hashlib.sha256(b"\x00" * 1024).digest()
    assert c.is_satisfied()
    assert c.num_cons == 1 << 19 and c.num_vars == 1 << 19
