"""Field-by-field comparison of two abi.BatchResult objects (oracle vs CUDA path / emulation)."""
import numpy as np

REC_FIELDS = ["mapped", "tid", "pos", "strand", "mapq", "nm", "x0", "x1", "xt", "n_alts", "n_hits", "best_lower", "best_lower_rev",
              "best_size", "absolute_pos", "frames_popped", "d_ext_steps", "lf_steps"]


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def compare_read(a, b, i, check_hits=True, check_counters=True, label_a="oracle", label_b="gpu"):
    """Raises AssertionError describing the first difference of read i.  Bit-exact on everything, including f32 scores
    (compared by bit pattern)."""
    ra, rb = a.records[i], b.records[i]
    for f in REC_FIELDS:
        if not check_counters and f in ("frames_popped", "d_ext_steps", "lf_steps"):
            continue
        if not ra["mapped"] and f in ("tid", "pos", "xt"):
            continue
        assert int(ra[f]) == int(rb[f]), "read %d field %s: %s=%s %s=%s" % (i, f, label_a, ra[f], label_b, rb[f])
    if ra["mapped"]:
        assert bits(ra["alignment_score"]) == bits(rb["alignment_score"]), "read %d AS" % i
        if ra["x1"] > 0:
            assert bits(ra["xs"]) == bits(rb["xs"]), "read %d XS" % i
        assert a.cigar_str(ra["cigar_off"], ra["cigar_len"]) == b.cigar_str(rb["cigar_off"], rb["cigar_len"]), "read %d CIGAR" % i
        assert a.md_str(ra["md_off"], ra["md_len"]) == b.md_str(rb["md_off"], rb["md_len"]), "read %d MD" % i
        for k in range(int(ra["n_alts"])):
            xa, xb = ra["alts"][k], rb["alts"][k]
            for f in ("tid", "strand", "pos", "nm", "interval_size"):
                assert int(xa[f]) == int(xb[f]), "read %d alt %d %s" % (i, k, f)
            assert bits(xa["alignment_score"]) == bits(xb["alignment_score"])
            assert a.cigar_str(xa["cigar_off"], xa["cigar_len"]) == b.cigar_str(xb["cigar_off"], xb["cigar_len"])
            assert a.md_str(xa["md_off"], xa["md_len"]) == b.md_str(xb["md_off"], xb["md_len"])
    if check_hits:
        ha, hb = a.hits_of(i), b.hits_of(i)
        assert len(ha) == len(hb), "read %d n_hits" % i
        for k, (x, y) in enumerate(zip(ha, hb)):
            assert (x["lower"], x["lower_rev"], x["size"]) == (y["lower"], y["lower_rev"], y["size"]), "read %d hit %d interval" % (i, k)
            assert bits(x["score"]) == bits(y["score"]), "read %d hit %d score" % (i, k)
            assert x["ops"] == y["ops"], "read %d hit %d edit ops\n%s\n%s" % (i, k, x["ops"], y["ops"])


def compare_results(a, b, check_hits=True, check_counters=True, label_a="oracle", label_b="gpu", n=None):
    """Raises AssertionError with a description of the first difference (first n reads if given)."""
    if n is None:
        assert len(a) == len(b), (len(a), len(b))
        n = len(a)
    for i in range(n):
        compare_read(a, b, i, check_hits, check_counters, label_a, label_b)
    if n == len(a) == len(b) and a.xa is not None and b.xa is not None:
        assert a.xa == b.xa
    return True


def mismatching_reads(a, b, n=None, check_hits=False, check_counters=True):
    """Indices (and first message) of the reads among the first n whose records differ — bench.py's in-run parity count."""
    n = min(len(a), len(b)) if n is None else n
    bad = []
    for i in range(n):
        try:
            compare_read(a, b, i, check_hits, check_counters)
        except AssertionError as e:
            bad.append((i, str(e)))
    return bad
