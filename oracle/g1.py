"""BN254 G1 — big-int restatement (TEST INFRASTRUCTURE ONLY).

The reference wraps `halo2curves::bn256::G1` (bn254/src/curve.rs:59-66,74), a
crates.io dependency (`halo2curves = "0.9"`, bn254/Cargo.toml:22) that is NOT in
/root/reference.  Its published definition is restated here: the curve
y^2 = x^3 + 3 over Fq, generator (1, 2), prime group order = Fr modulus.
Group elements are mathematically unique, so affine coordinates are canonical.

Points are affine tuples (x, y) of canonical ints, identity = None.
Wire format (what crosses the FFI): [x:4xu64][y:4xu64] little-endian Montgomery
limbs with R_q = 2^256 mod q; identity = all-zero (halo2curves G1Affine::identity).
"""
import numpy as np

from . import fr

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
B = 3
RQ = (1 << 256) % Q
RQ_INV = pow(RQ, -1, Q)
ORDER = fr.P
G = (1, 2)
MASK64 = (1 << 64) - 1


def is_on_curve(p):
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - B) % Q == 0


def neg(p):
    if p is None:
        return None
    return (p[0], (-p[1]) % Q)


def add(p1, p2):
    """Complete affine addition (P+P, P+(-P), identity inputs)."""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


# Jacobian internals for speed -----------------------------------------------------
def _jdbl(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A = X * X % Q
    Bv = Y * Y % Q
    C = Bv * Bv % Q
    D = 2 * ((X + Bv) * (X + Bv) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jadd(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jdbl(p)
        return (1, 1, 0)
    H = (U2 - U1) % Q
    Rr = (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


def _to_j(p):
    return (1, 1, 0) if p is None else (p[0], p[1], 1)


def _from_j(p):
    X, Y, Z = p
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def mul(p, k):
    """Scalar multiplication k*p, k any int (reduced mod group order)."""
    k %= ORDER
    acc = (1, 1, 0)
    base = _to_j(p)
    while k:
        if k & 1:
            acc = _jadd(acc, base)
        base = _jdbl(base)
        k >>= 1
    return _from_j(acc)


def msm(points, scalars):
    """G1::multi_exp (bn254/src/curve.rs:158-180): sum scalars[i]*points[i];
    empty -> identity; length mismatch -> assertion (the reference panics)."""
    assert len(points) == len(scalars), "points and scalars must have the same length"
    acc = (1, 1, 0)
    for p, s in zip(points, scalars):
        s %= ORDER
        if p is None or s == 0:
            continue
        acc = _jadd(acc, _to_j(mul(p, s)))
    return _from_j(acc)


def msm_via_dlog(dlogs, scalars):
    """Algebraic shortcut for a synthetic SRS P_i = dlogs[i]*G:
    MSM = (sum c_i * s_i mod r) * G.  O(n) field ops + one scalar mul."""
    acc = 0
    for d, s in zip(dlogs, scalars):
        acc = (acc + d * s) % ORDER
    return mul(G, acc)


# wire ---------------------------------------------------------------------------
def _limbs(v):
    return [(v >> (64 * i)) & MASK64 for i in range(4)]


def to_wire(points):
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            continue
        out[i, 0:4] = _limbs(p[0] * RQ % Q)
        out[i, 4:8] = _limbs(p[1] * RQ % Q)
    return out


def from_wire(arr):
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in a:
        xm = sum(int(row[k]) << (64 * k) for k in range(4))
        ym = sum(int(row[4 + k]) << (64 * k) for k in range(4))
        assert xm < Q and ym < Q, "non-canonical Fq limbs"
        if xm == 0 and ym == 0:
            out.append(None)
        else:
            p = (xm * RQ_INV % Q, ym * RQ_INV % Q)
            assert is_on_curve(p), "point not on curve"
            out.append(p)
    return out


# compressed form -------------------------------------------------------------------
# G1::to_bytes (bn254/src/curve.rs:136-139) = halo2curves G1Affine::to_bytes, Serialize / Deserialize for G1
# (:84-98).  halo2curves is not in /root/reference; its GroupEncoding for bn256 (two spare bits in the top byte
# of x) is restated from its published definition.  PARITY UNPINNED at the byte level: the reference holds no
# byte vector for any point, so the two layouts halo2curves has used are both provided.
ENC_HALO2 = 0   # 0.4 and later ("0.9" in bn254/Cargo.toml:22): sign = byte 31 bit 6, identity = byte 31 bit 7
ENC_LEGACY = 1  # 0.3 and earlier: sign = byte 31 bit 7, identity = 32 zero bytes


def to_bytes(p, enc=ENC_HALO2):
    if p is None:
        b = bytearray(32)
        if enc == ENC_HALO2:
            b[31] |= 0x80
        return bytes(b)
    x, y = p
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << (6 if enc == ENC_HALO2 else 7)
    return bytes(b)


def from_bytes(data, enc=ENC_HALO2):
    """Returns the point, or raises ValueError("Invalid G1 point") (bn254/src/curve.rs:95)."""
    b = bytearray(data)
    assert len(b) == 32
    if enc == ENC_HALO2:
        inf, sign = b[31] >> 7, (b[31] >> 6) & 1
        b[31] &= 0x3F
    else:
        inf, sign = 0, b[31] >> 7
        b[31] &= 0x7F
    x = int.from_bytes(b, "little")
    if inf:
        if x or sign:
            raise ValueError("Invalid G1 point")
        return None
    if enc == ENC_LEGACY and x == 0 and sign == 0:
        return None
    if x >= Q:
        raise ValueError("Invalid G1 point")
    rhs = (x * x * x + B) % Q
    y = pow(rhs, (Q + 1) // 4, Q)
    if y * y % Q != rhs:
        raise ValueError("Invalid G1 point")
    if (y & 1) != sign:
        y = (-y) % Q
    return (x, y)
