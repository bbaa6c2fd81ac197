"""Generates tests/golden/kzg_small.npz from the big-int oracle (oracle/), which is itself pinned
to the reference's known-answer tests (tests/test_oracle.py).  The reference is Rust and cannot be
built in this image, so these are oracle-generated fixtures, not outputs of the reference binary.

    python tests/golden/make_golden.py

Shapes follow the reference's own tests: DFT conformance h = 16, w = 3, shift = GENERATOR
(field-testing/src/dft_testing.rs:9-112); KZG with SRS 1024-style alpha = 12345
(eon-uni-stark/tests/fib_air.rs:112-136) on an 8 x 2 trace, opened at zeta and zeta*omega.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dft, fr, g1, kzg  # noqa: E402


def main():
    rng = np.random.default_rng(1)
    out = {}
    # ---- DFT family, h = 16, w = 3 --------------------------------------------------------
    h, w = 16, 3
    x_w = fr.random_wire(rng, h * w).reshape(h, w, 4)
    x = dft.mat_from_wire(x_w)
    s = fr.GENERATOR
    out["dft_in"] = x_w
    out["dft_out"] = dft.mat_to_wire(dft.dft_batch(x))
    out["coset_dft_out"] = dft.mat_to_wire(dft.coset_dft_batch(x, s))
    out["idft_out"] = dft.mat_to_wire(dft.idft_batch(x))
    out["coset_idft_out"] = dft.mat_to_wire(dft.coset_idft_batch(x, s))
    out["coset_lde1_out"] = dft.mat_to_wire(dft.coset_lde_batch(x, 1, s))
    out["lde2_out"] = dft.mat_to_wire(dft.lde_batch(x, 2))
    # ---- KZG, 8 x 2, alpha = 12345 ---------------------------------------------------------
    alpha = 12345
    srs = kzg.init_srs_unsafe(15, alpha)
    out["srs"] = g1.to_wire(srs)
    ev_w = fr.random_wire(rng, 8 * 2).reshape(8, 2, 4)
    ev = dft.mat_from_wire(ev_w)
    dom = (1, 3)
    commits, pdata = kzg.commit(srs, [(dom, ev)], fast=False)
    out["kzg_evals"] = ev_w
    out["kzg_coeffs"] = dft.mat_to_wire(pdata[0]["coeffs"])
    out["kzg_commit"] = g1.to_wire(commits[0])
    zeta = fr.from_wire(fr.random_wire(rng, 1))[0]
    zeta_next = zeta * fr.two_adic_generator(3) % fr.P
    opened, wits = kzg.open_(srs, [(pdata, [[zeta, zeta_next]])])
    out["kzg_points"] = fr.to_wire([zeta, zeta_next])
    out["kzg_opened"] = np.stack([fr.to_wire(v) for v in opened[0][0]])
    out["kzg_witness"] = np.stack([g1.to_wire(p) for p in wits[0][0]])
    qdom = (fr.GENERATOR, 4)
    out["kzg_evals_on_quotient_domain"] = dft.mat_to_wire(kzg.get_evaluations_on_domain(pdata[0], qdom))
    # shifted-domain commit (quotient chunk style): shift = 5 * omega_16
    sh = fr.GENERATOR * fr.two_adic_generator(4) % fr.P
    commits2, pdata2 = kzg.commit(srs, [((sh, 3), ev)], fast=False)
    out["kzg_shift"] = fr.to_wire([sh])
    out["kzg_commit_shifted"] = g1.to_wire(commits2[0])
    # compressed encodings (G1::to_bytes): the first SRS powers, the commitments, the identity and -G
    pts = list(srs[:6]) + list(commits[0]) + [None, g1.neg(g1.G)]
    out["g1_points"] = g1.to_wire(pts)
    out["g1_bytes_halo2"] = np.frombuffer(b"".join(g1.to_bytes(p, g1.ENC_HALO2) for p in pts), dtype=np.uint8).reshape(-1, 32)
    out["g1_bytes_legacy"] = np.frombuffer(b"".join(g1.to_bytes(p, g1.ENC_LEGACY) for p in pts), dtype=np.uint8).reshape(-1, 32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kzg_small.npz"), **out)
    print("wrote kzg_small.npz:", {k: v.shape for k, v in out.items()})


def batched():
    """tests/golden/kzg_batched.npz: Pcs::commit_quotient (commit/src/pcs.rs:82-102) of a 16 x 2 quotient on
    GENERATOR * <omega_16> in 2 chunks, and ONE KzgPcs::open (kzg/src/pcs.rs:289-335) over a trace round (8 x 2 at
    zeta, zeta*omega and 8 x 1 at zeta) and the chunk round (both chunks at zeta), alpha = 12345 — laid out flat in
    [round][matrix][point][column] order, the order eon_kzg_open_batch returns."""
    rng = np.random.default_rng(2)
    alpha = 12345
    srs = kzg.init_srs_unsafe(15, alpha)
    out = {"alpha": np.array([alpha], dtype=np.uint64)}
    t0_w = fr.random_wire(rng, 8 * 2).reshape(8, 2, 4)
    t1_w = fr.random_wire(rng, 8 * 1).reshape(8, 1, 4)
    q_w = fr.random_wire(rng, 16 * 2).reshape(16, 2, 4)
    out["trace0"], out["trace1"], out["quotient"] = t0_w, t1_w, q_w
    dom = (1, 3)
    c_t, pd_t = kzg.commit(srs, [(dom, dft.mat_from_wire(t0_w)), (dom, dft.mat_from_wire(t1_w))], fast=False)
    qdom = (fr.GENERATOR, 4)
    doms = kzg.split_domains(qdom, 2)
    subs = kzg.split_evals(2, dft.mat_from_wire(q_w))
    c_q, pd_q = kzg.commit(srs, list(zip(doms, subs)), fast=False)
    out["trace_commit0"], out["trace_commit1"] = g1.to_wire(c_t[0]), g1.to_wire(c_t[1])
    out["chunk_shifts"] = fr.to_wire([d[0] for d in doms])
    out["chunk_commits"] = np.stack([g1.to_wire(c) for c in c_q])
    out["chunk_coeffs"] = np.stack([dft.mat_to_wire(m["coeffs"]) for m in pd_q])
    zeta = fr.from_wire(fr.random_wire(rng, 1))[0]
    zeta_next = zeta * fr.two_adic_generator(3) % fr.P
    out["points"] = fr.to_wire([zeta, zeta_next])
    rounds = [(pd_t, [[zeta, zeta_next], [zeta]]), (pd_q, [[zeta], [zeta]])]
    opened, wits = kzg.open_(srs, rounds)
    flat_v, flat_w = [], []
    for r in range(len(rounds)):
        for m in range(len(opened[r])):
            for pnt in range(len(opened[r][m])):
                flat_v.extend(opened[r][m][pnt])
                flat_w.extend(wits[r][m][pnt])
    out["opened_flat"] = fr.to_wire(flat_v)
    out["witness_flat"] = g1.to_wire(flat_w)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kzg_batched.npz"), **out)
    print("wrote kzg_batched.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if "--batched-only" in sys.argv:
        batched()
    else:
        main()
        batched()
