"""CPU test of the host mirror's domain logic (plonky3_eon_b200/pcs.py: TwoAdicMultiplicativeCoset) against the
definitions it follows: field/src/coset.rs:55-90,241-245 and commit/src/domain.rs:131-221, and against the
oracle's restatement.  Pure host code: no GPU, no native library calls."""
import numpy as np
import pytest

from oracle import dft as odft
from oracle import fr, kzg as okzg
from plonky3_eon_b200 import field
from plonky3_eon_b200.pcs import GpuKzgPcs, TwoAdicMultiplicativeCoset as Coset

P = fr.P


def points(c):
    out, x = [], c.first_point()
    for _ in range(c.size()):
        out.append(x)
        x = c.next_point(x)
    return out


def test_field_constants_match_the_oracle():
    assert field.P == fr.P and field.GENERATOR == fr.GENERATOR and field.TWO_ADICITY == fr.TWO_ADICITY
    for bits in (0, 1, 5, 20, 28):
        g = field.two_adic_generator(bits)
        assert g == fr.two_adic_generator(bits)
        assert pow(g, 1 << bits, P) == 1 and (bits == 0 or pow(g, 1 << (bits - 1), P) != 1)
    x = 0x1234567890ABCDEF1234567890ABCDEF
    assert field.from_wire(field.to_wire(x)) == x
    assert (field.to_wire(x) == fr.to_wire([x])[0]).all()


def test_coset_points_and_generator():
    c = Coset(5, 4)
    pts = points(c)
    assert pts == okzg.coset_points(5, 4)
    assert len(set(pts)) == 16 and c.next_point(pts[-1]) == pts[0]   # closes after |H| steps
    with pytest.raises(ValueError):                                     # TwoAdicMultiplicativeCoset::new -> None
        Coset(0, 3)
    with pytest.raises(ValueError):
        Coset(1, 29)                                                    # log_size > TWO_ADICITY


@pytest.mark.parametrize("min_size,log", [(1, 0), (2, 1), (3, 2), (16, 4), (17, 5)])
def test_create_disjoint_domain(min_size, log):
    c = Coset(7, 3)
    d = c.create_disjoint_domain(min_size)                              # shift * GENERATOR, log2_ceil(min_size)
    assert (d.shift, d.log_size) == (7 * fr.GENERATOR % P, log)
    assert not set(points(c)) & set(points(d))                          # the cosets are disjoint (domain.rs:155-168)


@pytest.mark.parametrize("log_size,num_chunks", [(3, 1), (4, 2), (5, 4), (3, 8)])
def test_split_domains_and_evals_partition_the_coset(log_size, num_chunks):
    c = Coset(fr.GENERATOR, log_size)
    subs = c.split_domains(num_chunks)
    assert [(s.shift, s.log_size) for s in subs] == okzg.split_domains((fr.GENERATOR, log_size), num_chunks)
    # chunk i's j-th point is point i + j * num_chunks of the parent (domain.rs:174-186)
    all_pts = points(c)
    for i, s in enumerate(subs):
        assert points(s) == all_pts[i::num_chunks]
    # and split_evals deals the rows out the same way (domain.rs:188-221)
    rng = np.random.default_rng(log_size)
    ev = fr.random_wire(rng, (1 << log_size) * 2).reshape(1 << log_size, 2, 4)
    parts = c.split_evals(num_chunks, ev)
    want = okzg.split_evals(num_chunks, odft.mat_from_wire(ev))
    for a, b in zip(parts, want):
        assert odft.mat_from_wire(a) == b
    with pytest.raises(Exception):                                      # log2_strict_usize panics (domain.rs:175)
        c.split_domains(3)


@pytest.mark.parametrize("degree,log", [(0, 0), (1, 0), (2, 1), (5, 3), (8, 3), (9, 4), (1 << 20, 20)])
def test_natural_domain_for_degree(degree, log):
    pcs = GpuKzgPcs.__new__(GpuKzgPcs)                                  # host-only method: no context needed
    d = pcs.natural_domain_for_degree(degree)                           # kzg/src/pcs.rs:218-221
    assert (d.shift, d.log_size) == (1, log)
    assert (d.shift, d.log_size) == okzg.natural_domain_for_degree(degree)
