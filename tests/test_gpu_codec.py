"""GPU parity: compressed G1 points (G1::to_bytes, Serialize / Deserialize for G1 — bn254/src/curve.rs:84-98,
136-139) through eon_g1_compress / eon_g1_decompress / eon_srs_load_compressed vs the oracle's restatement of
halo2curves' GroupEncoding, byte for byte; and the CanObserve<KzgCommitment> element stream
(kzg/src/pcs.rs:409-438)."""
import os

import numpy as np
import pytest

from oracle import dft as odft
from oracle import fr, g1, kzg as okzg

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kzg_small.npz")
ENCS = [g1.ENC_HALO2, g1.ENC_LEGACY]


@pytest.fixture(scope="module")
def ctx():
    from plonky3_eon_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _bytes(points, enc):
    return np.frombuffer(b"".join(g1.to_bytes(p, enc) for p in points), dtype=np.uint8).reshape(-1, 32)


@pytest.mark.parametrize("enc", ENCS)
def test_golden_bytes(ctx, enc):
    g = np.load(GOLD)
    key = "g1_bytes_halo2" if enc == g1.ENC_HALO2 else "g1_bytes_legacy"
    assert (ctx.g1_to_bytes(g["g1_points"], enc) == g[key]).all()
    assert (ctx.g1_from_bytes(g[key], enc) == g["g1_points"]).all()


@pytest.mark.parametrize("enc", ENCS)
def test_compress_decompress_vs_oracle(ctx, enc):
    rng = np.random.default_rng(3)
    pts = [g1.G, g1.neg(g1.G), None, g1.mul(g1.G, fr.P - 1), None]
    pts += [g1.mul(g1.G, int.from_bytes(rng.bytes(32), "little")) for _ in range(200)]
    wire = g1.to_wire(pts)
    want = _bytes(pts, enc)
    got = ctx.g1_to_bytes(wire, enc)
    assert got.shape == (len(pts), 32) and (got == want).all()
    back = ctx.g1_from_bytes(want, enc)
    assert (back == wire).all()
    # the other sign bit decodes to the negated point
    flipped = want.copy()
    nz = np.array([p is not None for p in pts])
    flipped[nz, 31] ^= 0x40 if enc == g1.ENC_HALO2 else 0x80
    assert g1.from_wire(ctx.g1_from_bytes(flipped, enc)) == [g1.neg(p) for p in pts]
    # empty batch
    assert ctx.g1_to_bytes(np.zeros((0, 8), np.uint64), enc).shape == (0, 32)
    assert ctx.g1_from_bytes(np.zeros((0, 32), np.uint8), enc).shape == (0, 8)


@pytest.mark.parametrize("enc", ENCS)
def test_invalid_encodings_are_rejected_with_their_index(ctx, enc):
    from plonky3_eon_b200 import InvalidG1Point
    good = _bytes([g1.mul(g1.G, k) for k in range(1, 40)], enc)
    bads = [bytes([4] + [0] * 31),                  # x^3 + 3 is not a square
            g1.Q.to_bytes(32, "little"),            # x = q: not canonical
            (g1.Q + 1).to_bytes(32, "little")]      # x = q + 1 (would alias x = 1)
    if enc == g1.ENC_HALO2:
        bads += [bytes([1] + [0] * 30 + [0x80]),    # identity flag with x bits
                 bytes(31) + b"\xc0"]               # identity flag with the sign bit
    for k, bad in enumerate(bads):
        with pytest.raises(ValueError):
            g1.from_bytes(bad, enc)
        data = good.copy()
        pos = 5 + 3 * k
        data[pos] = np.frombuffer(bad, dtype=np.uint8)
        data[30] = np.frombuffer(bads[0], dtype=np.uint8)   # a later bad one: the first index is reported
        with pytest.raises(InvalidG1Point) as ei:
            ctx.g1_from_bytes(data, enc)
        assert ei.value.index == pos
    ctx.g1_from_bytes(good, enc)  # the context stays usable


@pytest.mark.parametrize("enc", ENCS)
def test_srs_from_serialised_bytes_commits_identically(ctx, enc):
    """A deserialised SRS (g1_powers as compressed bytes) gives the same resident points and the same
    commitments / openings as the SRS generated on the device."""
    from plonky3_eon_b200 import GpuKzgPcs, InvalidG1Point, TwoAdicMultiplicativeCoset, observe_commitment
    alpha, h, w = 12345, 64, 3
    srs = okzg.init_srs_unsafe(h - 1, alpha)
    ser = _bytes(srs, enc)
    pcs = GpuKzgPcs.from_srs_bytes(ser, ctx=ctx, enc=enc)
    assert pcs.max_degree == h - 1
    assert (pcs.g1_powers() == g1.to_wire(srs)).all()
    assert (pcs.g1_powers_bytes(enc=enc) == ser).all()
    rng = np.random.default_rng(8)
    evw = fr.random_wire(rng, h * w).reshape(h, w, 4)
    dom = TwoAdicMultiplicativeCoset(1, 6)
    commit, pdata = pcs.commit([(dom, evw)])
    ocommit, opdata = okzg.commit(srs, [((1, 6), odft.mat_from_wire(evw))])
    assert g1.from_wire(commit[0]) == ocommit[0]
    zeta = 0xDEADBEEF12345
    opened, proof = pcs.open([(pdata, [[zeta]])])
    oopened, owits = okzg.open_(srs, [(opdata, [[zeta]])])
    assert fr.from_wire(opened[0][0][0]) == oopened[0][0][0]
    assert g1.from_wire(proof[0][0][0]) == owits[0][0][0]
    # CanObserve<KzgCommitment>: 32 compressed bytes -> four LE u64 -> Fr::from_u64 (pcs.rs:409-438)
    want = []
    for p in ocommit[0]:
        b = g1.to_bytes(p, enc)
        want += [int.from_bytes(b[8 * i:8 * i + 8], "little") for i in range(4)]
    assert observe_commitment(commit, ctx=ctx, enc=enc) == want
    pdata[0].free()
    # a corrupt serialisation is refused and leaves no SRS behind
    bad = ser.copy()
    bad[17] = np.frombuffer(bytes([4] + [0] * 31), dtype=np.uint8)
    with pytest.raises(InvalidG1Point) as ei:
        GpuKzgPcs.from_srs_bytes(bad, ctx=ctx, enc=enc)
    assert ei.value.index == 17 and ctx.srs_size() == 0


def test_large_batch_roundtrip(ctx):
    """2^18 SRS powers: compress -> decompress is the identity, and every encoding has x < q with the flag bits
    consistent with y's parity (size-independent property; the oracle checks a sample)."""
    from plonky3_eon_b200 import GpuKzgPcs
    n = 1 << 18
    pcs = GpuKzgPcs.new(n - 1, 7, ctx=ctx)
    wire = pcs.g1_powers()
    b = ctx.g1_to_bytes(wire)
    assert (ctx.g1_from_bytes(b) == wire).all()
    assert not (b[:, 31] & 0x80).any()
    for i in (0, 1, 2, 12345, n - 1):
        assert g1.to_bytes(g1.from_wire(wire[i])[0]) == b[i].tobytes()
