"""Synthetic workloads of BASELINE.md / SURVEY.md §8d: i.i.d. genomes and simulated damaged reads.

Vectorised numpy generators with documented seeds (genome seed 42, reads seed 1000 + config number).
Used by bench.py and by the full-size parity checks; nothing here touches the GPU.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b

# name -> (genome bp, n contigs, read length range, library, reads in the full config)
CONFIGS = {
    "cfg1": dict(genome_bp=1_000_000, n_contigs=1, len_range=(50, 50), library="single_stranded", n_reads=100_000, seed=1001,
                 desc="1 Mbp random reference, 50 bp single_stranded reads, -p 0.03"),
    "cfg2": dict(genome_bp=4_600_000, n_contigs=1, len_range=(30, 75), library="double_stranded", n_reads=1_000_000, seed=1002,
                 desc="4.6 Mbp (E. coli-size) reference, 30-75 bp double_stranded reads, -p 0.03"),
    "cfg3": dict(genome_bp=50_000_000, n_contigs=8, len_range=(50, 50), library="single_stranded", n_reads=10_000_000, seed=1003,
                 desc="50 Mbp (chr21-scale) reference, 8 contigs, 50 bp single_stranded reads, -p 0.03"),
    "cfg4": dict(genome_bp=3_100_000_000, n_contigs=24, len_range=(25, 100), library="single_stranded", n_reads=100_000_000, seed=1004,
                 desc="3.1 Gbp hg19-scale reference, 24 contigs, 25-100 bp single_stranded reads, -p 0.03"),
}


def random_genome_array(n_bp, seed=42):
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=n_bp, dtype=np.uint8)]


def split_contigs(genome, n_contigs):
    """Equal-size contigs chr1..chrN over one genome array -> list[(name, bytes)]"""
    n = len(genome)
    cuts = [n * i // n_contigs for i in range(n_contigs + 1)]
    return [("chr%d" % (i + 1), genome[cuts[i]:cuts[i + 1]].tobytes()) for i in range(n_contigs)]


def simulate_batch(genome, n_reads, len_range, seed, library="single_stranded", exo_frac=0.10, f=0.5, t=0.5, d=0.02, s=1.0,
                   divergence=0.02, indel_rate=0.001):
    """Vectorised read simulator -> (seq u8, qual u8, offsets u64) packed arrays.
    Endogenous reads: uniform start, random strand, divergence substitutions, one indel with probability
    indel_rate * L, deamination drawn from the model's own C->T (and, double-stranded, G->A) probabilities,
    then sequencing errors at 10^(-q/10).  10 % exogenous (i.i.d. random) reads."""
    rng = np.random.default_rng(seed)
    G = len(genome)
    lo, hi = len_range
    L = rng.integers(lo, hi + 1, size=n_reads)
    Lmax = int(hi) + 1
    start = (rng.random(n_reads) * (G - Lmax - 1)).astype(np.int64)
    cols = np.arange(Lmax)
    m = genome[start[:, None] + cols[None, :]]  # n x Lmax window (one spare column for deletions)
    valid = cols[None, :] < L[:, None]
    # strand
    rev = rng.random(n_reads) < 0.5
    idx = np.where(rev[:, None], (L[:, None] - 1 - cols[None, :]) % Lmax, cols[None, :])
    m = np.take_along_axis(m, idx, axis=1)
    m[rev] = _COMP[m[rev]]
    # divergence
    mut = (rng.random(m.shape) < divergence) & valid
    m[mut] = ACGT[rng.integers(0, 4, size=int(mut.sum()))]
    # indels: one per affected read, away from the ends
    ind = (rng.random(n_reads) < indel_rate * L) & (L > 20)
    for r in np.nonzero(ind)[0]:
        p = int(rng.integers(8, L[r] - 8))
        row = m[r, : L[r]].copy()
        if rng.random() < 0.5:
            row = np.delete(row, p)
        else:
            row = np.insert(row, p, ACGT[int(rng.integers(0, 4))])[: Lmax]
        L[r] = len(row)
        m[r, : len(row)] = row
    valid = cols[None, :] < L[:, None]
    # deamination
    i = cols[None, :].astype(np.float64)
    pf = f ** (i + 1.0)
    pt = t ** (L[:, None] - i)
    u = rng.random(m.shape)
    if library == "single_stranded":
        p_fwd = pf + pt - pf * pt
        p_c = s * p_fwd + d * (1 - p_fwd)
        m[(m == ord("C")) & (u < p_c) & valid] = ord("T")
    else:
        p_c = s * pf + d * (1 - pf)
        p_g = s * pt + d * (1 - pt)
        dc = (m == ord("C")) & (u < p_c) & valid
        dg = (m == ord("G")) & (u < p_g) & valid
        m[dc] = ord("T")
        m[dg] = ord("A")
    # exogenous reads
    exo = rng.random(n_reads) < exo_frac
    m[exo] = ACGT[rng.integers(0, 4, size=(int(exo.sum()), Lmax))]
    # qualities + sequencing error
    qv = np.array([40, 30, 20, 2], dtype=np.uint8)
    q = qv[rng.choice(4, size=m.shape, p=[0.70, 0.20, 0.08, 0.02])]
    err = (rng.random(m.shape) < 10.0 ** (-q.astype(np.float64) / 10.0)) & valid
    m[err] = ACGT[rng.integers(0, 4, size=int(err.sum()))]
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    np.cumsum(L, out=offsets[1:])
    return m[valid].copy(), q[valid].copy(), offsets


def algorithmic_bytes(records, total_bases):
    """128 B per popped frame, 128 B per D-array extension step, 64 B per LF step,
    8 B per located position (folded into W), 2 B per base in, 64 B per result out."""
    P = int(records["frames_popped"].astype(np.int64).sum())
    E = int(records["d_ext_steps"].astype(np.int64).sum())
    W = int(records["lf_steps"].astype(np.int64).sum())
    n = len(records)
    total = 128 * P + 128 * E + 64 * W + 2 * int(total_bases) + 64 * n
    return dict(P=P, E=E, W=W, search_bytes=128 * P, darray_bytes=128 * E, locate_bytes=64 * W, total_bytes=total)
