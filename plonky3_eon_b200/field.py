"""Host-side Fr helpers for the wire format (no compute path: constants + int<->limb packing).

Fr wire = 4 x u64 little-endian Montgomery limbs (bn254/src/field.rs:96-105).
"""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % P
R_INV = pow(R, -1, P)
GENERATOR = 5            # Fr::GENERATOR, bn254/src/field.rs:372-377
TWO_ADICITY = 28         # bn254/src/field.rs:564
_OMEGA28 = pow(GENERATOR, (P - 1) >> TWO_ADICITY, P)
_M64 = (1 << 64) - 1


def two_adic_generator(bits):
    """bn254/src/field.rs:567-573."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(_OMEGA28, 1 << (TWO_ADICITY - bits), P)


def to_wire(x):
    """canonical int -> np.uint64[4] Montgomery limbs."""
    m = (int(x) % P) * R % P
    return np.array([(m >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


def from_wire(w):
    w = np.asarray(w, dtype=np.uint64).reshape(4)
    m = sum(int(w[i]) << (64 * i) for i in range(4))
    return m * R_INV % P


def log2_strict(n):
    """p3_util::log2_strict_usize (util/src/lib.rs:39): panics on non powers of two."""
    if n <= 0 or n & (n - 1):
        raise ValueError(f"Not a power of two: {n}")
    return n.bit_length() - 1
