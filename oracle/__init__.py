"""CPU oracle for the BN254 KZG hot path of Lolazyx/plonky3-eon.

TEST INFRASTRUCTURE ONLY.  Nothing under ``plonky3_eon_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker.

Parity status: the reference is Rust and its curve/MSM arithmetic lives in the
un-vendored crate ``halo2curves = "0.9"`` (bn254/Cargo.toml:22, no Cargo.lock),
so the reference cannot be built or run in this image.  The oracle is a
restatement pinned against every known-answer value the reference's own tests
and constants hold for this path:
Fr Montgomery constants (bn254/src/field.rs:29-53,256-281,372-377,556-561),
NaiveDft::basic (dft/src/naive.rs:49-85), divide_by_height / coset_shift_cols
KATs (dft/src/util.rs:49-138), the MSM identities of bn254/src/curve.rs:597-628
and the KZG vectors of kzg/src/tests.rs:19-171.  There are no byte-level golden
outputs anywhere in the reference ("parity unpinned" at the byte level); because
every output on this path is a mathematically unique value (canonical Fr limbs,
affine G1 points), the algebraic KATs pin the bytes.
"""
from . import fr, dft, g1, kzg  # noqa: F401
