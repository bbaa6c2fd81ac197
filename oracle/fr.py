"""BN254 scalar field Fr — big-int restatement (TEST INFRASTRUCTURE ONLY).

Follows bn254/src/field.rs and bn254/src/helpers.rs of the reference:
  * P, MU, R^2                      field.rs:29-53
  * monty_mul (4 x IMR rounds)      helpers.rs:168-205
  * add / sub                       field.rs:464-508
  * two_adic_generator              field.rs:556-574
  * uniform sampler                 field.rs:534-551

Two representations are used:
  * "canonical" Python ints in [0, P)       (what the maths is done in)
  * "wire"      4 x u64 little-endian Montgomery limbs, a*R mod P, R = 2^256
                (what crosses the FFI; Fr.value in the reference)
"""
import numpy as np

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = (1 << 256) % P
R_INV = pow(R, -1, P)
R2 = (R * R) % P
MU64 = pow(P, -1, 1 << 64)          # +P^{-1} mod 2^64 (the reduction subtracts), field.rs:40
MASK64 = (1 << 64) - 1
TWO_ADICITY = 28
GENERATOR = 5                        # field.rs:372-377
# 5^((P-1)/2^28), field.rs:553-561 (doc comment value is canonical; limbs there are Montgomery)
TWO_ADIC_GENERATOR = pow(GENERATOR, (P - 1) >> TWO_ADICITY, P)


def limbs(x):
    """256-bit int -> [u64;4] little-endian."""
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def from_limbs(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def to_mont(a):
    """canonical -> Montgomery integer (a*R mod P)."""
    return (a * R) % P


def from_mont(m):
    return (m * R_INV) % P


# --- literal restatement of the reference's 64-bit interleaved Montgomery reduction ----
def _imr(acc0, acc):
    """interleaved_monty_reduction, helpers.rs:168-179.  acc0 = low limb, acc = upper 256 bits."""
    t = (acc0 * MU64) & MASK64
    u = (P * t) >> 64                     # mul_small(P, t) upper 4 limbs
    sub = acc - u
    if sub < 0:
        sub += P
    return sub


def monty_mul_limbs(lhs, rhs):
    """monty_mul, helpers.rs:188-205, on Montgomery *integers* (lhs < P)."""
    assert lhs < P
    r = limbs(rhs)
    x = lhs * r[0]
    res = _imr(x & MASK64, x >> 64)
    for i in (1, 2, 3):
        x = lhs * r[i] + res
        res = _imr(x & MASK64, x >> 64)
    return res


def mont_mul(a_m, b_m):
    """Montgomery product on Montgomery integers; equals a_m*b_m*R^{-1} mod P."""
    return (a_m * b_m * R_INV) % P


def add(a, b):
    return (a + b) % P


def sub(a, b):
    return (a - b) % P


def mul(a, b):
    return (a * b) % P


def inv(a):
    return pow(a, -1, P)


def two_adic_generator(bits):
    """field.rs:567-573: omega_28 squared (28-bits) times."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(TWO_ADIC_GENERATOR, 1 << (TWO_ADICITY - bits), P)


# --- wire format helpers (numpy uint64, Montgomery) --------------------------------
def to_wire(vals):
    """iterable of canonical ints -> np.uint64 array [n,4] of Montgomery limbs."""
    vals = list(vals)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (v * R) % P
        out[i, 0] = m & MASK64
        out[i, 1] = (m >> 64) & MASK64
        out[i, 2] = (m >> 128) & MASK64
        out[i, 3] = m >> 192
    return out


def from_wire(arr):
    """np.uint64 [...,4] Montgomery limbs -> list of canonical ints (asserts < P)."""
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a:
        m = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        assert m < P, "non-canonical Montgomery limbs"
        out.append((m * R_INV) % P)
    return out


def random_wire(rng, n):
    """Reference sampler (field.rs:534-551): 32 random bytes, clear top 2 bits, reject >= P,
    use AS the Montgomery limbs.  rng: np.random.Generator.  Returns np.uint64 [n,4]."""
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    pl = limbs(P)
    while filled < n:
        need = n - filled
        cand = rng.integers(0, 1 << 64, size=(need + 8, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 62) - 1)
        # lexicographic compare from the top limb
        lt = np.zeros(len(cand), dtype=bool)
        eq = np.ones(len(cand), dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (cand[:, k] < np.uint64(pl[k]))
            eq &= cand[:, k] == np.uint64(pl[k])
        good = cand[lt][:need]
        out[filled:filled + len(good)] = good
        filled += len(good)
    return out
