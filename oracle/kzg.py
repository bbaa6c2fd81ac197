"""KZG commit / open / evaluate-on-domain — big-int restatement (TEST INFRASTRUCTURE ONLY).

Follows kzg/src/params.rs:123-139 (init_srs_unsafe), kzg/src/util.rs:37-40
(commit_column), :63-68 (eval_poly), :100-111 (quotient_and_eval) and
kzg/src/pcs.rs:218-335 (natural_domain_for_degree, commit,
get_evaluations_on_domain, open) of the reference.
Matrices: lists of rows of canonical ints.  Points: oracle.g1 affine tuples.
"""
from . import dft, fr, g1

P = fr.P


class DegreeTooLarge(Exception):
    """KzgError::DegreeTooLarge (kzg/src/params.rs:178-211)."""


def srs_dlogs(max_degree, alpha):
    """Discrete logs alpha^i of g1_powers[i] (init_srs_unsafe, params.rs:129-132)."""
    out = []
    p = 1
    for _ in range(max_degree + 1):
        out.append(p)
        p = p * alpha % P
    return out


def init_srs_unsafe(max_degree, alpha):
    """params.rs:123-139, G1 part only (g2_alpha is verifier-side, out of scope)."""
    return [g1.mul(g1.G, d) for d in srs_dlogs(max_degree, alpha)]


def commit_column(srs, coeffs):
    """util.rs:37-40."""
    deg = max(len(coeffs) - 1, 0)
    if deg > len(srs) - 1:
        raise DegreeTooLarge(f"degree {deg} > max {len(srs) - 1}")
    return g1.msm(srs[:len(coeffs)], coeffs)


def eval_poly(coeffs, point):
    """util.rs:63-68 (Horner from the top)."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * point + c) % P
    return acc


def quotient_and_eval(coeffs, point):
    """util.rs:100-111 (synthetic division)."""
    if not coeffs:
        return [], 0
    quotient = [0] * (len(coeffs) - 1)
    carry = coeffs[-1]
    for i in range(len(coeffs) - 2, -1, -1):
        quotient[i] = carry
        carry = (coeffs[i] + carry * point) % P
    return quotient, carry


def natural_domain_for_degree(degree):
    """pcs.rs:218-221 -> (shift, log_size)."""
    npow = 1
    while npow < degree:
        npow <<= 1
    return (1, dft.log2_strict(npow))


def commit(srs, matrices, fast=True):
    """pcs.rs:223-265.  matrices: list of ((shift, log_size), evals).
    Returns (commitments[matrix][col], prover_data[matrix] = dict(domain, evals, coeffs))."""
    commitments, prover = [], []
    for (shift, log_size), evals in matrices:
        h = len(evals)
        w = len(evals[0]) if h else 0
        assert h == 1 << log_size, "evaluation height must match domain size"
        if max(h - 1, 0) > len(srs) - 1:
            raise DegreeTooLarge(f"degree {h - 1} > max {len(srs) - 1}")
        coeffs = dft.coset_idft_batch(evals, shift, fast)
        cols = [commit_column(srs, [coeffs[r][c] for r in range(h)]) for c in range(w)]
        commitments.append(cols)
        prover.append({"domain": (shift, log_size), "evals": evals, "coeffs": coeffs})
    return commitments, prover


def coset_points(shift, log_size):
    """field/src/coset.rs:241-245: shift * g^i in natural order."""
    g = fr.two_adic_generator(log_size)
    out, x = [], shift % P
    for _ in range(1 << log_size):
        out.append(x)
        x = x * g % P
    return out


def get_evaluations_on_domain(prover_matrix, domain):
    """pcs.rs:267-287 — the quadratic Horner evaluation, verbatim semantics."""
    if prover_matrix["domain"] == tuple(domain):
        return [row[:] for row in prover_matrix["evals"]]
    coeffs = prover_matrix["coeffs"]
    h = len(coeffs)
    w = len(coeffs[0]) if h else 0
    out = []
    for point in coset_points(*domain):
        out.append([eval_poly([coeffs[r][c] for r in range(h)], point) for c in range(w)])
    return out


def open_(srs, rounds):
    """pcs.rs:289-335.  rounds: list of (prover_data, points_per_matrix).
    Returns (opened_values[round][matrix][point][col], witnesses[round][matrix][point][col])."""
    opened, proofs = [], []
    for prover_data, points_per_matrix in rounds:
        assert len(prover_data) == len(points_per_matrix)
        mv, mp = [], []
        for matrix, points in zip(prover_data, points_per_matrix):
            coeffs = matrix["coeffs"]
            h = len(coeffs)
            w = len(coeffs[0]) if h else 0
            vals_m, wit_m = [], []
            for z in points:
                evals, wits = [], []
                for c in range(w):
                    q, v = quotient_and_eval([coeffs[r][c] for r in range(h)], z)
                    evals.append(v)
                    wits.append(commit_column(srs, q))
                vals_m.append(evals)
                wit_m.append(wits)
            mv.append(vals_m)
            mp.append(wit_m)
        opened.append(mv)
        proofs.append(mp)
    return opened, proofs


# commit_quotient default: commit/src/pcs.rs:82-102 + commit/src/domain.rs:174-221 ------
def split_domains(domain, num_chunks):
    shift, log_size = domain
    lc = dft.log2_strict(num_chunks)
    g = fr.two_adic_generator(log_size)
    return [(shift * pow(g, i, P) % P, log_size - lc) for i in range(num_chunks)]


def split_evals(num_chunks, evals):
    return [[row[:] for row in evals[i::num_chunks]] for i in range(num_chunks)]


# ---- KzgMmcs (kzg/src/mmcs.rs) ------------------------------------------------------------------
def _log2_ceil(n):
    """n.next_power_of_two().trailing_zeros() (mmcs.rs:203,209)."""
    lg = 0
    while (1 << lg) < n:
        lg += 1
    return lg


def mmcs_commit(srs, matrices):
    """mmcs.rs:155-190: every column of every matrix, taken as coefficients, through commit_column.
    Returns commitments[matrix][col]."""
    out = []
    for m in matrices:
        h = len(m)
        deg = max(h - 1, 0)
        if deg > len(srs) - 1:
            raise DegreeTooLarge(f"degree {deg} > max {len(srs) - 1}")
        w = len(m[0]) if h else 0
        out.append([commit_column(srs, [m[r][c] for r in range(h)]) for c in range(w)])
    return out


def mmcs_local_index(index, height, log2_max_height):
    """mmcs.rs:209-214."""
    lg = _log2_ceil(height)
    li = index >> (log2_max_height - lg) if log2_max_height >= lg else index
    return li % height


def mmcs_open_batch(srs, index, matrices):
    """mmcs.rs:192-237.  Returns (opened_values[matrix][col], witnesses[matrix][col])."""
    max_height = max((len(m) for m in matrices), default=0)
    lmax = _log2_ceil(max_height)
    opened, wits = [], []
    for m in matrices:
        h = len(m)
        w = len(m[0]) if h else 0
        point = mmcs_local_index(index, h, lmax)
        row, mw = [], []
        for c in range(w):
            q, v = quotient_and_eval([m[r][c] for r in range(h)], point)
            row.append(v)
            mw.append(commit_column(srs, q))
        opened.append(row)
        wits.append(mw)
    return opened, wits
