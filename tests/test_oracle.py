"""Pin the oracle against every KAT the reference holds for this path."""
import numpy as np
import pytest

import os

from oracle import dft, fr, g1, kzg

P = fr.P
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_small.npz")


def L(x):
    return fr.limbs(x)


# ---- Fr constants: bn254/src/field.rs:29-53,256-281,372-377,556-561 -------------------
def test_fr_constants():
    assert L(P) == [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029]
    assert fr.MU64 == 0x3d1e0a6c10000001
    assert L(fr.R2) == [0x1bb8e645ae216da7, 0x53fe3ab1e35c59e3, 0x8c49833d53bb8085, 0x0216d0b17f4e44a5]
    assert L(fr.to_mont(1)) == [0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f]
    assert L(fr.to_mont(2)) == [0x592c68389ffffff6, 0x6df8ed2b3ec19a53, 0xccdd46def0f28c5c, 0x1c14ef83340fbe5e]
    assert L(fr.to_mont(P - 1)) == [0x974bc177a0000006, 0xf13771b2da58a367, 0x51e1a2470908122e, 0x2259d6b14729c0fa]
    assert L(fr.to_mont(5)) == [0x1b0d0ef99fffffe6, 0xeaba68a3a32a913f, 0x47d8eb76d8dd0689, 0x15d0085520f5bbc3]
    assert L(fr.to_mont(fr.TWO_ADIC_GENERATOR)) == [
        0x636e735580d13d9c, 0xa22bf3742445ffd6, 0x56452ac01eb203d8, 0x1860ef942963f9e7]
    assert fr.TWO_ADIC_GENERATOR == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert fr.to_mont(1) == 6350874878119819312338956282401532410528162663560392320966563075034087161851
    assert fr.to_mont(5) == 9866131518759821339448375666750386964092448917385927261134611188594627313638


def test_two_adicity():
    # field.rs:635-648: 2-adicity 28; omega_28 has exact order 2^28
    assert (P - 1) % (1 << 28) == 0 and ((P - 1) >> 28) % 2 == 1
    w = fr.two_adic_generator(28)
    assert pow(w, 1 << 28, P) == 1 and pow(w, 1 << 27, P) == P - 1
    assert fr.two_adic_generator(0) == 1 and fr.two_adic_generator(1) == P - 1
    # SURVEY App. B
    w8 = fr.two_adic_generator(3)
    assert w8 == 0x2b337de1c8c14f22ec9b9e2f96afef3652627366f8170a0a948dad4ac1bd5e80
    assert pow(w8, 4, P) == P - 1
    assert fr.inv(8) == 0x2a57c4a4850b6c2481463cffb1512d51832d6b3f6a82427f1b65b6e172000001


def test_monty_mul_restatement_matches_definition():
    rng = np.random.default_rng(1)
    w = fr.random_wire(rng, 64)
    ms = [fr.from_limbs(r) for r in w]
    for a, b in zip(ms[:32], ms[32:]):
        assert fr.monty_mul_limbs(a, b) == fr.mont_mul(a, b)
    # rhs unconstrained (helpers.rs:185): as_canonical uses rhs = [1,0,0,0]
    for a in ms[:8]:
        assert fr.monty_mul_limbs(a, 1) == fr.from_mont(a)
    # new(): monty_mul(R^2, [v,0,0,0]) (field.rs:110-116)
    assert fr.monty_mul_limbs(fr.R2, 12345) == fr.to_mont(12345)


def test_wire_roundtrip_and_sampler():
    rng = np.random.default_rng(7)
    w = fr.random_wire(rng, 100)
    vals = fr.from_wire(w)
    assert all(0 <= v < P for v in vals)
    assert np.array_equal(fr.to_wire(vals), w)


# ---- DFT: dft/src/naive.rs:49-104, dft/src/util.rs:49-138 ----------------------------
def test_naive_basic():
    mat = [[5, 2, 0], [4, 3, 0]]
    assert dft.naive_dft_batch(mat) == [[9, 5, 0], [1, P - 1, 0]]


def test_divide_by_height_and_shift():
    assert dft.divide_by_height([[2, 4], [6, 8]]) == [[1, 2], [3, 4]]
    assert dft.divide_by_height([[7, 9]]) == [[7, 9]]
    with pytest.raises(ValueError):
        dft.divide_by_height([[1], [2], [3]])
    assert dft.coset_shift_cols([[1, 2], [3, 4], [5, 6]], 2) == [[1, 2], [6, 8], [20, 24]]


@pytest.mark.parametrize("log_h", range(0, 5))
def test_fast_matches_naive_and_roundtrips(log_h):
    # field-testing/src/dft_testing.rs:9-112 shapes: h = 1..16, w = 3, shift = GENERATOR
    rng = np.random.default_rng(1)
    h, w = 1 << log_h, 3
    mat = dft.mat_from_wire(fr.random_wire(rng, h * w).reshape(h, w, 4))
    assert dft.fast_dft_batch(mat) == dft.naive_dft_batch(mat)
    s = fr.GENERATOR
    assert dft.idft_batch(dft.dft_batch(mat)) == mat
    assert dft.coset_idft_batch(dft.coset_dft_batch(mat, s), s) == mat
    assert dft.coset_idft_batch(dft.coset_dft_batch(mat, s, True), s, True) == mat
    # coset_lde: row j of the output = column polynomials evaluated at s*omega_{2h}^j
    lde = dft.coset_lde_batch(mat, 1, s)
    coeffs = dft.idft_batch(mat)
    pts = kzg.coset_points(s, log_h + 1)
    for j in range(2 * h):
        for c in range(w):
            assert lde[j][c] == kzg.eval_poly([coeffs[r][c] for r in range(h)], pts[j])
    assert dft.coset_lde_batch(mat, 1, s, True) == lde


# ---- G1: bn254/src/curve.rs:523-533,597-628; SURVEY App. B ---------------------------
KG = {
    2: (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
        0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4),
    5: (0x17c139df0efee0f766bc0204762b774362e4ded88953a39ce849a8a7fa163fa9,
        0x01e0559bacb160664764a357af8a9fe70baa9258e0b959273ffc5718c6d4cc7c),
    8: (0x08b1d51d23480c10f472f5e93b9cfea88238c121fe155af7043937882c306a63,
        0x299836713dad3fa34e337aa412466015c366af8ec50b9d7bd05aa74642822021),
    76: (0x02a63beba6b22ff50c1a46ec2368713e1e1ed5e413732ddf571760800e102718,
         0x00406bb65ad052a53c8c47c9ccc29e59c585e0894c54805081b8fa5e3925a285),
    5377: (0x1b3f932b6c4da2949f2c5f1e07e60afec1596972f642ef651e10554a175aa94d,
           0x303f2679044d3b96803fe456e8f3e6c28b4f8a6fd139db41018f2ba47e0515e9),
    134: (0x1fa72e4cd19ec67c0abb4b44e0a37a6f5b6375edd3941dcafae4589da797ba60,
          0x0566b4cd3ac681682567aa52e47a6d870fd5f0a77cde426f3f3f6dfdbe699390),
    19703: (0x187100ff57101144e5549944871d38d3821a2612d30331eb553a95012dfba689,
            0x29245abf6b6b713eaabc81dd8d881dd73bce996291a0b1f6eff11a251af1e731),
    502: (0x18aedecb55ba9abc8591d6ed19dd947a3456a39286a6866ef38809839d6b23fe,
          0x090ff2212b557b5c9aae1f57d971b7406b18612a2f5b7918f89f587a99aecc57),
    1999: (0x12c933cc2979a273a7356580079bdd5ddaea9c048eab4dedf402789e917b58c5,
           0x067e405cada4cf71e138de4af82f4f8481894fbe61061f04ea1754e6c8126041),
}


def test_g1_multiples_of_generator():
    for k, pt in KG.items():
        assert g1.is_on_curve(pt)
        assert g1.mul(g1.G, k) == pt
    # group order
    assert g1.mul(g1.G, g1.ORDER) is None
    assert g1.add(g1.G, g1.neg(g1.G)) is None
    assert g1.add(g1.G, g1.G) == KG[2]


def test_g1_multi_exp_kats():
    # curve.rs:597-628
    assert g1.msm([], []) is None
    assert g1.msm([g1.G], [5]) == KG[5]
    assert g1.msm([g1.G, g1.G], [2, 3]) == KG[5]
    assert g1.msm([g1.mul(g1.G, 7), g1.mul(g1.G, 11)], [3, 5]) == KG[76]
    with pytest.raises(AssertionError):
        g1.msm([g1.G], [1, 2])


def test_g1_wire_montgomery_example():
    w = g1.to_wire([KG[2], None])
    x = sum(int(w[0, k]) << (64 * k) for k in range(4))
    y = sum(int(w[0, 4 + k]) << (64 * k) for k in range(4))
    assert x == 0x13227397098d014dc2822db40c0ac2ecbc0b548b438e5469e10460b6c3e7ea38
    assert y == 0x04644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
    assert not w[1].any()
    assert g1.from_wire(w) == [KG[2], None]


# ---- KZG: kzg/src/tests.rs:19-171 ---------------------------------------------------
def test_kzg_batch_verification_vectors():
    srs = kzg.init_srs_unsafe(16, 42)
    assert kzg.commit_column(srs, [1, 2, 3]) == KG[5377]
    assert kzg.commit_column(srs, [5, 7, 11]) == KG[19703]
    q, v = kzg.quotient_and_eval([1, 2, 3], 2)
    assert (q, v) == ([8, 3], 17)
    assert kzg.commit_column(srs, q) == KG[134]
    q, v = kzg.quotient_and_eval([5, 7, 11], 3)
    assert (q, v) == ([40, 11], 125)
    assert kzg.commit_column(srs, q) == KG[502]


def test_kzg_single_vector():
    srs = kzg.init_srs_unsafe(8, 999)
    assert kzg.commit_column(srs, [1, 2]) == KG[1999]
    q, v = kzg.quotient_and_eval([1, 2], 5)
    assert (q, v) == ([2], 11)
    assert kzg.commit_column(srs, q) == KG[2]


def test_pcs_roundtrip_vector():
    # kzg/src/tests.rs:19-48: alpha = 7, evals x+1 on the size-8 subgroup
    srs = kzg.init_srs_unsafe(8, 7)
    dom = (1, 3)
    evals = [[(x + 1) % P] for x in kzg.coset_points(*dom)]
    commits, pdata = kzg.commit(srs, [(dom, evals)])
    assert pdata[0]["coeffs"] == [[1], [1]] + [[0]] * 6
    assert commits[0][0] == KG[8]
    opened, wit = kzg.open_(srs, [(pdata, [[2]])])
    assert opened[0][0][0] == [3]
    assert wit[0][0][0][0] == g1.G
    # degree guard (params.rs:164-173)
    with pytest.raises(kzg.DegreeTooLarge):
        kzg.commit_column(srs, [1] * 10)


def test_empty_and_one_row():
    srs = kzg.init_srs_unsafe(4, 3)
    assert kzg.quotient_and_eval([], 5) == ([], 0)
    assert kzg.quotient_and_eval([9], 5) == ([], 9)
    assert kzg.commit_column(srs, []) is None
    commits, pdata = kzg.commit(srs, [((1, 0), [[4, 5]])])
    assert commits[0] == [g1.mul(g1.G, 4), g1.mul(g1.G, 5)]
    opened, wit = kzg.open_(srs, [(pdata, [[11]])])
    assert opened[0][0][0] == [4, 5] and wit[0][0][0] == [None, None]


def test_msm_dlog_shortcut_and_evals_on_domain():
    rng = np.random.default_rng(3)
    n = 16
    dl = kzg.srs_dlogs(n - 1, 12345)
    srs = [g1.mul(g1.G, d) for d in dl]
    sc = fr.from_wire(fr.random_wire(rng, n))
    assert g1.msm(srs, sc) == g1.msm_via_dlog(dl, sc)
    # get_evaluations_on_domain == zero-pad + coset DFT
    evals = [[s] for s in sc[:8]]
    _, pdata = kzg.commit(srs, [((1, 3), evals)])
    dom = (fr.GENERATOR, 4)
    quad = kzg.get_evaluations_on_domain(pdata[0], dom)
    assert quad == dft.coset_lde_batch(evals, 1, fr.GENERATOR)
    assert kzg.get_evaluations_on_domain(pdata[0], (1, 3)) == evals


def test_split_domains_and_evals():
    dom = (5, 3)
    subs = kzg.split_domains(dom, 2)
    g = fr.two_adic_generator(3)
    assert subs == [(5, 2), (5 * g % P, 2)]
    ev = [[i] for i in range(8)]
    assert kzg.split_evals(2, ev) == [[[0], [2], [4], [6]], [[1], [3], [5], [7]]]


# ---- KzgMmcs: kzg/src/tests.rs:50-70 (mmcs_roundtrip) and the doc example of mmcs.rs:36-56 ------
def test_mmcs_roundtrip_values():
    """alpha = 5, max_degree 4, matrix [[1,2],[3,4]]: columns are the polynomials 1+3X and 2+4X.
    commit = (1+3*5) G, (2+4*5) G; open_batch(0): values (1, 2), quotients [3], [4] -> witnesses 3G, 4G.
    The reference checks these through the pairing (verify_batch); here the same group elements are
    pinned directly (they are the unique elements that make e(C - vG, H) = e(W, (alpha - z) H) hold)."""
    from oracle import g1, kzg
    srs = kzg.init_srs_unsafe(4, 5)
    m = [[1, 2], [3, 4]]
    com = kzg.mmcs_commit(srs, [m])
    assert com[0][0] == g1.mul(g1.G, 16) and com[0][1] == g1.mul(g1.G, 22)
    opened, wits = kzg.mmcs_open_batch(srs, 0, [m])
    assert opened == [[1, 2]]
    assert wits[0] == [g1.mul(g1.G, 3), g1.mul(g1.G, 4)]
    # KZG opening identity in the exponent: (f(alpha) - f(z)) = q(alpha) * (alpha - z), all rows
    for idx in range(2):
        opened, wits = kzg.mmcs_open_batch(srs, idx, [m])
        for c, f_alpha in enumerate((16, 22)):
            q_alpha = (f_alpha - opened[0][c]) * pow(5 - idx, -1, kzg.P) % kzg.P
            assert wits[0][c] == g1.mul(g1.G, q_alpha)


def test_mmcs_local_index_mixed_heights():
    """mmcs.rs:203-214: index is scaled down to shorter matrices; non-power-of-two heights wrap."""
    from oracle import kzg
    assert kzg.mmcs_local_index(5, 8, 3) == 5
    assert kzg.mmcs_local_index(5, 4, 3) == 2
    assert kzg.mmcs_local_index(5, 2, 3) == 1
    assert kzg.mmcs_local_index(7, 6, 3) == 1      # log2_ceil(6) = 3 -> 7 % 6
    assert kzg.mmcs_local_index(3, 1, 3) == 0


# ---- compressed G1 encoding (bn254/src/curve.rs:84-98,136-139 -> halo2curves GroupEncoding) ------------
@pytest.mark.parametrize("enc", [g1.ENC_HALO2, g1.ENC_LEGACY])
def test_g1_compressed_known_points(enc):
    sign_bit = 0x40 if enc == g1.ENC_HALO2 else 0x80
    # generator (1, 2): x = 1, y even
    assert g1.to_bytes(g1.G, enc) == bytes([1] + [0] * 31)
    # -G = (1, q - 2): q is odd, so y is odd -> sign bit set
    b = g1.to_bytes(g1.neg(g1.G), enc)
    assert b[0] == 1 and b[31] == sign_bit and not any(b[1:31])
    ident = g1.to_bytes(None, enc)
    assert ident == (bytes(31) + b"\x80" if enc == g1.ENC_HALO2 else bytes(32))
    for p in (g1.G, g1.neg(g1.G), None, g1.mul(g1.G, 5377), g1.mul(g1.G, fr.P - 1)):
        assert g1.from_bytes(g1.to_bytes(p, enc), enc) == p


@pytest.mark.parametrize("enc", [g1.ENC_HALO2, g1.ENC_LEGACY])
def test_g1_compressed_roundtrip_and_rejects(enc):
    rng = np.random.default_rng(77)
    for _ in range(40):
        p = g1.mul(g1.G, int.from_bytes(rng.bytes(32), "little"))
        b = g1.to_bytes(p, enc)
        assert len(b) == 32 and g1.from_bytes(b, enc) == p
        # the other sign bit gives the negated point
        flip = bytearray(b)
        flip[31] ^= 0x40 if enc == g1.ENC_HALO2 else 0x80
        assert g1.from_bytes(bytes(flip), enc) == g1.neg(p)
    # x = 4: 4^3 + 3 = 67 is not a square mod q -> rejected; x >= q -> rejected
    assert pow(67, (g1.Q - 1) // 2, g1.Q) == g1.Q - 1
    with pytest.raises(ValueError):
        g1.from_bytes(bytes([4] + [0] * 31), enc)
    with pytest.raises(ValueError):
        g1.from_bytes(g1.Q.to_bytes(32, "little"), enc)
    if enc == g1.ENC_HALO2:
        with pytest.raises(ValueError):   # identity flag with other bits set
            g1.from_bytes(bytes([1] + [0] * 30 + [0x80]), enc)
        with pytest.raises(ValueError):
            g1.from_bytes(bytes(31) + b"\xc0", enc)


def test_g1_compressed_golden():
    g = np.load(GOLD)
    pts = g1.from_wire(g["g1_points"])
    for enc, key in ((g1.ENC_HALO2, "g1_bytes_halo2"), (g1.ENC_LEGACY, "g1_bytes_legacy")):
        for p, row in zip(pts, g[key]):
            assert g1.to_bytes(p, enc) == row.tobytes()
            assert g1.from_bytes(row.tobytes(), enc) == p
