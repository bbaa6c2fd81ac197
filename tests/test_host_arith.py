"""CPU check of the CUDA library's field/curve arithmetic (csrc/fp.cuh, ec.cuh) compiled for
the host with a software carry flag, against the big-int oracle.  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import fr, g1

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "hostlib.cpp")
LIB = os.path.join(HERE, "host", "libhosteon.so")


@pytest.fixture(scope="module")
def lib():
    csrc = os.path.join(HERE, "..", "plonky3_eon_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("fp.cuh", "ec.cuh", "consts.cuh", "fp_shoup.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def to_int(row):
    return sum(int(row[k]) << (64 * k) for k in range(4))


def from_int(v):
    return np.array([(v >> (64 * k)) & ((1 << 64) - 1) for k in range(4)], dtype=np.uint64)


def rand_mont(rng, mod, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    for i in range(n):
        v = int.from_bytes(rng.bytes(40), "little") % mod
        out[i] = from_int(v)
    return out


EDGE = lambda mod: [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (mod - 1) // 2, (1 << 253), mod >> 1]


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_field_ops(lib, which, mod):
    rng = np.random.default_rng(11 + which)
    edge = EDGE(mod)
    pairs = [(a, b) for a in edge for b in edge]
    a = np.concatenate([np.array([from_int(x) for x, _ in pairs]), rand_mont(rng, mod, 300)])
    b = np.concatenate([np.array([from_int(y) for _, y in pairs]), rand_mont(rng, mod, 300)])
    n = len(a)
    r = np.zeros_like(a)
    Rinv = pow(1 << 256, -1, mod)
    expect = {
        0: lambda x, y: x * y * Rinv % mod,
        1: lambda x, y: (x + y) % mod,
        2: lambda x, y: (x - y) % mod,
        3: lambda x, y: (-x) % mod,
        5: lambda x, y: x * Rinv % mod,
        6: lambda x, y: x * (1 << 256) % mod,
        7: lambda x, y: x * x * Rinv % mod,
    }
    for op, fn in expect.items():
        lib.host_fp_op(which, op, _ptr(a), _ptr(b), _ptr(r), n)
        for i in range(n):
            assert to_int(r[i]) == fn(to_int(a[i]), to_int(b[i])), (which, op, i)
    # inverse: a * inv(a) == R (Montgomery one); inv(0) == 0
    m = 24
    lib.host_fp_op(which, 4, _ptr(a[-m:]), _ptr(b[-m:]), _ptr(r[-m:]), m)
    for i in range(n - m, n):
        x = to_int(a[i])
        assert to_int(r[i]) * x * Rinv % mod == ((1 << 256) % mod if x else 0)


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_binary_gcd_inverse(lib, which, mod):
    """fp_inv (binary extended Euclid, csrc/fp.cuh) against pow(x, -1, p) and against the Fermat chain it replaced:
    edge values, powers of two (long runs of halvings), values next to the modulus and random elements."""
    rng = np.random.default_rng(2024 + which)
    R = (1 << 256) % mod
    vals = EDGE(mod) + [1 << k for k in range(0, 254, 7)] + [mod - (1 << k) for k in range(0, 250, 11)]
    vals += [(1 << k) - 1 for k in (2, 31, 32, 33, 64, 127, 128, 200, 253)] + [3, 5, mod // 3, R, (R * R) % mod]
    a = np.concatenate([np.array([from_int(v % mod) for v in vals]), rand_mont(rng, mod, 3000)])
    n = len(a)
    r, f = np.zeros_like(a), np.zeros_like(a)
    lib.host_fp_op(which, 10, _ptr(a), _ptr(a), _ptr(r), n)
    lib.host_fp_op(which, 8, _ptr(a[:64]), _ptr(a[:64]), _ptr(f[:64]), 64)
    assert np.array_equal(r[:64], f[:64])
    for i in range(n):
        x = to_int(a[i])                       # Montgomery form of x / R
        want = pow(x, -1, mod) * R * R % mod if x else 0
        assert to_int(r[i]) == want, (which, i, hex(x))


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_safegcd_inverse(lib, which, mod):
    """fp_inv_safegcd (Bernstein-Yang divsteps, 20 batches of 30 on signed 30-bit limbs; the inversion on the
    latency-critical paths of the MSM) against pow(x, -1, p) and against the binary GCD: zero, one, values next to
    the modulus, powers of two and their complements (long runs of even / odd steps), 20000 random elements."""
    rng = np.random.default_rng(77 + which)
    R = (1 << 256) % mod
    vals = EDGE(mod) + [1 << k for k in range(0, 254)] + [mod - (1 << k) for k in range(0, 253)]
    vals += [(1 << k) - 1 for k in range(2, 254, 3)] + [3, 5, mod // 2, mod // 2 + 1, mod // 3, R, (R * R) % mod]
    a = np.concatenate([np.array([from_int(v % mod) for v in vals]), rand_mont(rng, mod, 20000)])
    n = len(a)
    r, b = np.zeros_like(a), np.zeros_like(a)
    lib.host_fp_op(which, 9, _ptr(a), _ptr(a), _ptr(r), n)
    lib.host_fp_op(which, 10, _ptr(a), _ptr(a), _ptr(b), n)
    assert np.array_equal(r, b)
    for i in range(0, n, 7):
        x = to_int(a[i])
        want = pow(x, -1, mod) * R * R % mod if x else 0
        assert to_int(r[i]) == want, (which, i, hex(x))


def test_fr_mul_matches_reference_limb_algorithm(lib):
    rng = np.random.default_rng(5)
    a = fr.random_wire(rng, 64)
    b = fr.random_wire(rng, 64)
    r = np.zeros_like(a)
    lib.host_fp_op(0, 0, _ptr(a), _ptr(b), _ptr(r), 64)
    for i in range(64):
        assert to_int(r[i]) == fr.monty_mul_limbs(to_int(a[i]), to_int(b[i]))


def test_curve_ops(lib):
    G = g1.G
    pts = [g1.mul(G, k) for k in (1, 2, 3, 7, 11, 12345)]
    out = np.zeros(8, dtype=np.uint64)

    def gsum(lst):
        w = g1.to_wire(lst)
        lib.host_g1_sum(_ptr(w), len(lst), _ptr(out))
        return g1.from_wire(out)[0]

    assert gsum([]) is None
    assert gsum([G]) == G
    assert gsum([G, G]) == g1.mul(G, 2)                 # doubling branch
    assert gsum([G, g1.neg(G)]) is None                 # inverse branch
    assert gsum([G, g1.neg(G), pts[3]]) == pts[3]       # restart from identity
    assert gsum([None, G, None, G, G]) == g1.mul(G, 3)  # identity inputs
    assert gsum(pts) == g1.mul(G, 1 + 2 + 3 + 7 + 11 + 12345)
    # full addition incl. equal / opposite operands
    def full(a0, a1, b0, b1):
        ws = [g1.to_wire([p]) for p in (a0, a1, b0, b1)]
        lib.host_g1_add_full(*[_ptr(w) for w in ws], _ptr(out))
        return g1.from_wire(out)[0]
    assert full(pts[0], pts[1], pts[2], pts[3]) == g1.mul(G, 13)
    assert full(pts[0], pts[1], pts[1], pts[0]) == g1.mul(G, 6)            # A == B -> doubling
    assert full(pts[0], pts[1], g1.neg(pts[0]), g1.neg(pts[1])) is None     # A == -B
    assert full(pts[0], g1.neg(pts[0]), pts[2], pts[3]) == g1.mul(G, 10)    # A identity
    assert full(pts[2], pts[3], pts[0], g1.neg(pts[0])) == g1.mul(G, 10)    # B identity
    # scalar multiplication
    rng = np.random.default_rng(2)
    for k in [0, 1, 2, 5, fr.P - 1, fr.P, int.from_bytes(rng.bytes(31), "little")]:
        kw = from_int(k)
        pw = g1.to_wire([pts[4]])
        lib.host_g1_mul(_ptr(pw), _ptr(kw), _ptr(out))
        assert g1.from_wire(out)[0] == g1.mul(pts[4], k)
    for k in [0, 1, 3, 32767, 65535]:
        pw = g1.to_wire([pts[2]])
        lib.host_g1_mul_u32(_ptr(pw), k, _ptr(out))
        assert g1.from_wire(out)[0] == g1.mul(pts[2], 2 * k)


def test_lazy_ntt_butterflies(lib):
    """The lazily reduced butterflies of csrc/ntt.cu (inputs anywhere in [0, 4r) forward / [0, 2r)
    inverse, one conditional correction each) agree with the canonical definitions
    (a + t b, a - t b) and (u + v, (u - v) t) of dft/src/butterflies.rs:52-63,118-132 after the final
    canonicalisation, including the extreme representatives."""
    P = fr.P
    R = 1 << 256
    rng = np.random.default_rng(9)
    rinv = pow(R, -1, P)
    for dif in (0, 1):
        top = 2 * P if dif else 4 * P
        cases = [(0, 0), (top - 1, top - 1), (top - 1, 0), (0, top - 1), (P, P - 1), (2 * P - 1, 1)]
        cases += [(int.from_bytes(rng.bytes(40), "little") % top, int.from_bytes(rng.bytes(40), "little") % top)
                  for _ in range(200)]
        for a, b in cases:
            tw = int.from_bytes(rng.bytes(40), "little") % P        # canonical Montgomery twiddle
            A, B, T = from_int(a), from_int(b), from_int(tw)
            o0 = np.zeros(4, dtype=np.uint64)
            o1 = np.zeros(4, dtype=np.uint64)
            lib.host_lazy_butterfly(dif, _ptr(A), _ptr(B), _ptr(T), _ptr(o0), _ptr(o1))
            if not dif:
                t = tw * b * rinv % P                               # Montgomery product
                want = ((a + t) % P, (a - t) % P)
            else:
                want = ((a + b) % P, (a - b) * tw * rinv % P)
            assert (to_int(o0), to_int(o1)) == want, (dif, a, b)


def _wide_cases(rng, n_rand):
    """256-bit operands that stress the Karatsuba glue: equal / ordered halves, all-ones limbs, sparse limbs."""
    M = (1 << 256) - 1
    lo = (1 << 128) - 1
    specials = [0, 1, M, lo, lo << 128, 1 << 128, (1 << 128) - 1 + (1 << 255), 0x80000000 << 96,
                (0xFFFFFFFF << 224) | 1, sum(0xFFFFFFFF << (64 * k) for k in range(4)),
                sum(1 << (32 * k + 31) for k in range(8)), (5 << 128) | 7, (7 << 128) | 5, (9 << 128) | 9]
    vals = list(specials)
    for _ in range(n_rand):
        v = int.from_bytes(rng.bytes(32), "little")
        kind = rng.integers(0, 6)
        if kind == 1:      # halves equal
            v = (v & lo) | ((v & lo) << 128)
        elif kind == 2:    # halves differ only in the lowest limb
            v = (v & lo) | (((v & lo) ^ 1) << 128)
        elif kind == 3:    # runs of all-ones limbs (long carry chains)
            v |= 0xFFFFFFFFFFFFFFFF << (64 * int(rng.integers(0, 4)))
        elif kind == 4:    # sparse
            v &= sum(0xFFFFFFFF << (32 * int(k)) for k in rng.integers(0, 8, size=3))
        vals.append(v)
    return vals


def _from_int16(v):
    return np.array([(v >> (32 * k)) & 0xFFFFFFFF for k in range(16)], dtype=np.uint32)


def test_wide_product_and_square(lib):
    """detail::mul8_wide / sqr8_wide (Karatsuba level + dedicated square) against big-int products on
    full 256-bit operands, including every pair of the special patterns."""
    rng = np.random.default_rng(5)
    vals = _wide_cases(rng, 400)
    sp = vals[:14]
    pairs = [(x, y) for x in sp for y in sp] + list(zip(vals, reversed(vals)))
    a = np.array([from_int(x) for x, _ in pairs])
    b = np.array([from_int(y) for _, y in pairs])
    T = np.zeros((len(pairs), 16), dtype=np.uint32)
    lib.host_wide_product(0, _ptr(a), _ptr(b), _ptr(T), len(pairs))
    for i, (x, y) in enumerate(pairs):
        assert (T[i] == _from_int16(x * y)).all(), ("mul", hex(x), hex(y))
    lib.host_wide_product(1, _ptr(a), _ptr(b), _ptr(T), len(pairs))
    for i, (x, _) in enumerate(pairs):
        assert (T[i] == _from_int16(x * x)).all(), ("sqr", hex(x))


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_split_product_equals_word_serial(lib, which, mod):
    """The split Montgomery product (Karatsuba + separate reduction) and the dedicated square return the
    same limbs as the word-serial CIOS form, for a < p and b anywhere in [0, 2^256) (the lazily reduced
    NTT butterflies feed values up to 4p), and the value is a*b/R mod p in [0, 2p)."""
    rng = np.random.default_rng(21 + which)
    bs = _wide_cases(rng, 600)
    as_ = [v % mod for v in _wide_cases(rng, len(bs) - 14)]
    as_ = (as_ + EDGE(mod) * 2)[:len(bs)]
    a = np.array([from_int(x) for x in as_])
    b = np.array([from_int(y) for y in bs])
    n = len(bs)
    r0, r1, r2 = np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)
    lib.host_fp_mul_lazy(which, 0, _ptr(a), _ptr(b), _ptr(r0), n)
    lib.host_fp_mul_lazy(which, 1, _ptr(a), _ptr(b), _ptr(r1), n)
    lib.host_fp_mul_lazy(which, 2, _ptr(a), _ptr(b), _ptr(r2), n)
    Rinv = pow(1 << 256, -1, mod)
    for i in range(n):
        x, y = as_[i], bs[i]
        assert to_int(r1[i]) == to_int(r0[i]), (which, i, hex(x), hex(y))
        assert to_int(r1[i]) < 2 * mod and to_int(r1[i]) % mod == x * y * Rinv % mod
        assert to_int(r2[i]) < 2 * mod and to_int(r2[i]) % mod == x * x * Rinv % mod


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_shoup_fixed_operand_product(lib, which, mod):
    """fp_shoup.cuh: r = a*w - q~*p with q~ from the truncated high product of a and wq = floor(w 2^256 / p).
    Checked limb for limb against the same truncated formula in big ints, and: r == a*w (mod p), r < 3p,
    for a anywhere in [0, 2^256) (all-ones limbs, 4p-1, ...) and w over random and edge twiddles."""
    rng = np.random.default_rng(77 + which)
    as_ = _wide_cases(rng, 700)
    ws = [v % mod for v in _wide_cases(rng, len(as_) - 18)]
    ws = (ws + EDGE(mod) * 2)[:len(as_)]
    n = len(as_)
    wqs = [(w << 256) // mod for w in ws]
    a = np.array([from_int(x) for x in as_])
    w = np.array([from_int(x) for x in ws])
    wq = np.array([from_int(x) for x in wqs])
    r = np.zeros_like(a)
    rc = np.zeros_like(a)
    lib.host_shoup_mul(which, 0, _ptr(a), _ptr(w), _ptr(wq), _ptr(r), n)
    lib.host_shoup_mul(which, 1, _ptr(a), _ptr(w), _ptr(wq), _ptr(rc), n)
    M32 = (1 << 32) - 1
    worst = 0
    for k in range(n):
        x, y, yq = as_[k], ws[k], wqs[k]
        xl = [(x >> (32 * i)) & M32 for i in range(8)]
        ql = [(yq >> (32 * i)) & M32 for i in range(8)]
        trunc = sum(xl[i] * ql[j] << (32 * (i + j)) for i in range(8) for j in range(8) if i + j >= 6)
        qt = trunc >> 256
        want = (x * y - qt * mod) % (1 << 256)
        got = to_int(r[k])
        assert got == want, (which, k, hex(x), hex(y))
        assert got < 3 * mod and got % mod == x * y % mod
        assert to_int(rc[k]) == x * y % mod
        worst = max(worst, got // mod)
    assert worst <= 2


@pytest.mark.parametrize("which,mod", [(0, fr.P), (1, g1.Q)])
def test_shoup_precompute(lib, which, mod):
    """(w, floor(w 2^256 / p)) from the Montgomery form of w: the exact division by p done as a multiplication
    by p^-1 mod 2^256."""
    rng = np.random.default_rng(5 + which)
    ws = [int.from_bytes(rng.bytes(40), "little") % mod for _ in range(300)] + EDGE(mod)
    rho = np.array([from_int(w * (1 << 256) % mod) for w in ws])
    w = np.zeros_like(rho)
    wq = np.zeros_like(rho)
    lib.host_shoup_precompute(which, _ptr(rho), _ptr(w), _ptr(wq), len(ws))
    for k, x in enumerate(ws):
        assert to_int(w[k]) == x
        assert to_int(wq[k]) == (x << 256) // mod


def test_lazy_ntt_butterflies_shoup_twiddles(lib):
    """The butterflies of k_ntt_pass with fixed-operand (Shoup) twiddles: same canonical results as the
    definitions, and the lazily reduced outputs stay inside the ranges the next layer assumes
    ([0, 4r) forward, [0, 2r) inverse), for inputs over the whole admissible range."""
    P = fr.P
    R = 1 << 256
    rng = np.random.default_rng(19)
    rinv = pow(R, -1, P)
    for dif in (0, 1):
        top = 2 * P if dif else 4 * P
        cases = [(0, 0), (top - 1, top - 1), (top - 1, 0), (0, top - 1), (P, P - 1), (2 * P - 1, 1), (top - 1, 1)]
        cases += [(int.from_bytes(rng.bytes(40), "little") % top, int.from_bytes(rng.bytes(40), "little") % top)
                  for _ in range(300)]
        tws = [0, 1, P - 1, R % P] + [int.from_bytes(rng.bytes(40), "little") % P for _ in range(len(cases))]
        for k, (a, b) in enumerate(cases):
            tw = tws[k % len(tws)]                                   # Montgomery form of the twiddle
            A, B, T = from_int(a), from_int(b), from_int(tw)
            o0, o1, r0, r1 = (np.zeros(4, dtype=np.uint64) for _ in range(4))
            lib.host_lazy_butterfly_shoup(dif, _ptr(A), _ptr(B), _ptr(T), _ptr(o0), _ptr(o1), _ptr(r0), _ptr(r1))
            if not dif:
                t = tw * b * rinv % P
                want = ((a + t) % P, (a - t) % P)
            else:
                want = ((a + b) % P, (a - b) * tw * rinv % P)
            assert (to_int(o0), to_int(o1)) == want, (dif, a, b, tw)
            assert to_int(r0) < top and to_int(r1) < top, (dif, a, b, tw)
            assert to_int(r0) % P == want[0] and to_int(r1) % P == want[1]
