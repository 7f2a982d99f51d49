"""Shared helpers for the test-suite: build oracle / product parameter objects from a case spec with
f32 arithmetic identical to the reference's (all derived values are computed in np.float32)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as ora  # noqa: E402  (test infrastructure)

f32 = np.float32


from mapad_b200.specs import model_div as _model_div, product_params, resolve_gap  # noqa: E402,F401  (oracle-free, shared with bench / tools)


def oracle_params(spec):
    """spec: dict(model=..., bound=..., gaps=...) as in tests/ref_cases.py -> OracleParams"""
    p = ora.OracleParams()
    model = _model_div(spec["model"])
    if model[0] == "test":
        p.model_test(model[1], model[2], model[3])
    elif model[0] == "vindija":
        p.model_vindija()
    else:
        p.model_simple(model[1], model[2], model[3], model[4], model[5], model[6], model[7])
    repr_mm = p.representative_mismatch_penalty()
    b = spec["bound"]
    if b[0] == "test":
        p.bound_test(b[1], repr_mm if b[2] is None else b[2])
    elif b[0] == "discrete":
        p.bound_discrete(b[1], b[2], repr_mm)
    else:
        p.bound_continuous(b[1], b[2], repr_mm)
    g = spec["gaps"]
    p.gaps(resolve_gap(g[0], repr_mm), resolve_gap(g[1], repr_mm), g[2], g[3], spec.get("abort", False))
    if "limits" in spec:
        p.limits(*spec["limits"])
    p.repr_mm = repr_mm
    return p


def revcomp(s):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    return "".join(comp.get(c, c) for c in reversed(s))


# -------------------------------------------------------------------------------------------------
# Synthetic genomes / reads: splitmix64-seeded numpy generators, documented seeds.
# -------------------------------------------------------------------------------------------------
def random_genome(n_bp, seed=42):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n_bp)].tobytes().decode()


def simulate_reads(genome, n_reads, len_range, seed, library="single_stranded", exo_frac=0.10, f=0.5, t=0.5, d=0.02, s=1.0,
                   divergence=0.02, indel_rate=0.001):
    """Damaged reads drawn from `genome` (both strands) plus `exo_frac` exogenous reads.
    Returns (list[bytes] seqs, list[bytes] quals)."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    g = np.frombuffer(genome.encode(), dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    qv = np.array([40, 30, 20, 2], dtype=np.uint8)
    qp = np.array([0.70, 0.20, 0.08, 0.02])
    seqs, quals = [], []
    for _ in range(n_reads):
        L = int(rng.integers(len_range[0], len_range[1] + 1))
        if rng.random() < exo_frac:
            r = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L)].copy()
        else:
            st = int(rng.integers(0, G - L + 1))
            r = g[st : st + L].copy()
            if rng.random() < 0.5:
                r = comp[r[::-1]]
            # divergence
            mut = rng.random(L) < divergence
            r[mut] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(mut.sum()))]
            # indels
            if rng.random() < indel_rate * L and L > 20:
                p = int(rng.integers(8, L - 8))
                if rng.random() < 0.5:
                    r = np.delete(r, p)
                else:
                    r = np.insert(r, p, b"ACGT"[int(rng.integers(0, 4))])
                L = len(r)
            # deamination
            i = np.arange(L)
            pf = f ** (i + 1.0)
            pt = t ** (L - i)
            if library == "single_stranded":
                p_fwd = pf + pt - pf * pt
                p_c = s * p_fwd + d * (1 - p_fwd)
                deam = (r == ord("C")) & (rng.random(L) < p_c)
                r[deam] = ord("T")
            else:
                p_c = s * pf + d * (1 - pf)
                p_g = s * pt + d * (1 - pt)
                dc = (r == ord("C")) & (rng.random(L) < p_c)
                dg = (r == ord("G")) & (rng.random(L) < p_g)
                r[dc] = ord("T")
                r[dg] = ord("A")
        q = qv[rng.choice(4, size=L, p=qp)]
        err = rng.random(L) < 10.0 ** (-q.astype(np.float64) / 10.0)
        r[err] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(err.sum()))]
        seqs.append(r.tobytes())
        quals.append(q.tobytes())
    return seqs, quals
