"""ctypes mirror of include/mapad_gpu.h (the C ABI of the hot path).

Every structure here must match the header byte for byte; tests/test_abi.py checks the sizes
against the values the shared library reports.
"""
import ctypes as C

import numpy as np

SDM_GET_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint8, C.c_uint8, C.c_uint8)
SDM_START_FN = C.CFUNCTYPE(C.c_int16, C.c_void_p, C.c_size_t)

MODEL_SIMPLE_ADNA, MODEL_VINDIJA_PWM, MODEL_TEST, MODEL_CUSTOM = 0, 1, 2, 3
LIB_SINGLE_STRANDED, LIB_DOUBLE_STRANDED = 0, 1
BOUND_CONTINUOUS, BOUND_DISCRETE, BOUND_TEST = 0, 1, 2
ED_INSERTION, ED_DELETION, ED_MATCH, ED_MISMATCH = 0, 1, 2, 3
BATCH_WANT_HITS, BATCH_RESIDENT, BATCH_NO_D2H, BATCH_UPLOAD_ONLY = 1, 2, 4, 8


class Params(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32),
        ("library", C.c_int32),
        ("five_prime_overhang", C.c_float),
        ("three_prime_overhang", C.c_float),
        ("ds_deamination_rate", C.c_float),
        ("ss_deamination_rate", C.c_float),
        ("divergence", C.c_float),
        ("ignore_base_quality", C.c_int32),
        ("test_deam_score", C.c_float),
        ("test_mm_score", C.c_float),
        ("test_match_score", C.c_float),
        ("custom_get", SDM_GET_FN),
        ("custom_start", SDM_START_FN),
        ("custom_user", C.c_void_p),
        ("bound_kind", C.c_int32),
        ("poisson_threshold", C.c_float),
        ("base_error_rate", C.c_float),
        ("cutoff", C.c_float),
        ("exponent", C.c_float),
        ("test_threshold", C.c_float),
        ("test_representative_mm", C.c_float),
        ("representative_mismatch_penalty", C.c_float),
        ("penalty_gap_open", C.c_float),
        ("penalty_gap_extend", C.c_float),
        ("gap_dist_ends", C.c_uint8),
        ("max_num_gaps_open", C.c_uint8),
        ("stack_limit_abort", C.c_uint8),
        ("reserved0", C.c_uint8),
        ("stack_limit", C.c_uint32),
        ("edit_tree_limit", C.c_uint32),
    ]


class IndexView(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("bwt", C.POINTER(C.c_uint8)),
        ("less", C.c_uint64 * 8),
        ("sentinel_rows", C.c_uint64 * 2),
        ("sa_sample", C.POINTER(C.c_uint64)),
        ("n_sa_samples", C.c_uint64),
        ("sa_rate", C.c_uint64),
        ("extra_rows", C.POINTER(C.c_uint64)),
        ("n_extra_rows", C.c_uint64),
        ("n_contigs", C.c_uint64),
        ("contig_start", C.POINTER(C.c_uint64)),
        ("contig_end", C.POINTER(C.c_uint64)),
        ("contig_name", C.POINTER(C.c_char_p)),
        ("orig_pos", C.POINTER(C.c_uint64)),
        ("orig_sym", C.POINTER(C.c_uint8)),
        ("n_orig", C.c_uint64),
    ]


class Reads(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("seq", C.c_void_p),
        ("qual", C.c_void_p),
        ("offsets", C.c_void_p),
        ("seeds", C.c_void_p),
        ("custom_penalties", C.c_void_p),
    ]


class EditOp(C.Structure):
    _fields_ = [("pos", C.c_uint16), ("kind", C.c_uint8), ("base", C.c_uint8)]


class Hit(C.Structure):
    _fields_ = [
        ("lower", C.c_uint64),
        ("lower_rev", C.c_uint64),
        ("size", C.c_uint64),
        ("alignment_score", C.c_float),
        ("edit_off", C.c_uint32),
        ("edit_len", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class Alt(C.Structure):
    _fields_ = [
        ("tid", C.c_int32),
        ("strand", C.c_int32),
        ("pos", C.c_int64),
        ("cigar_off", C.c_uint32),
        ("cigar_len", C.c_uint32),
        ("md_off", C.c_uint32),
        ("md_len", C.c_uint32),
        ("nm", C.c_int32),
        ("alignment_score", C.c_float),
        ("interval_size", C.c_uint64),
    ]


class Record(C.Structure):
    _fields_ = [
        ("mapped", C.c_int32),
        ("tid", C.c_int32),
        ("pos", C.c_int64),
        ("strand", C.c_int32),
        ("mapq", C.c_int32),
        ("alignment_score", C.c_float),
        ("nm", C.c_int32),
        ("x0", C.c_int32),
        ("x1", C.c_int32),
        ("xs", C.c_float),
        ("xt", C.c_int32),
        ("cigar_off", C.c_uint32),
        ("cigar_len", C.c_uint32),
        ("md_off", C.c_uint32),
        ("md_len", C.c_uint32),
        ("n_alts", C.c_uint32),
        ("alts", Alt * 2),
        ("hit_off", C.c_uint32),
        ("n_hits", C.c_uint32),
        ("best_lower", C.c_uint64),
        ("best_lower_rev", C.c_uint64),
        ("best_size", C.c_uint64),
        ("absolute_pos", C.c_uint64),
        ("frames_popped", C.c_uint32),
        ("d_ext_steps", C.c_uint32),
        ("lf_steps", C.c_uint32),
        ("flags", C.c_uint32),
    ]


class Results(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("records", C.POINTER(Record)),
        ("hits", C.POINTER(Hit)),
        ("n_hits", C.c_uint64),
        ("edit_ops", C.POINTER(EditOp)),
        ("n_edit_ops", C.c_uint64),
        ("cigar", C.POINTER(C.c_uint32)),
        ("n_cigar", C.c_uint64),
        ("text", C.POINTER(C.c_char)),
        ("n_text", C.c_uint64),
        ("ms_h2d", C.c_float),
        ("ms_prologue", C.c_float),
        ("ms_search", C.c_float),
        ("ms_epilogue", C.c_float),
        ("ms_d2h", C.c_float),
        ("ms_total", C.c_float),
        ("gpu_launches", C.c_uint64),
    ]


RECORD_DTYPE = np.dtype(Record)
HIT_DTYPE = np.dtype(Hit)
EDIT_OP_DTYPE = np.dtype(EditOp)


def _as_array(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (int(n) * dtype.itemsize)).from_address(C.addressof(ptr.contents))
    return np.frombuffer(buf, dtype=dtype).copy()


class BatchResult:
    """Owned (copied) numpy view of a mapad_results."""

    def __init__(self, res: Results):
        self.records = _as_array(res.records, res.n_reads, RECORD_DTYPE)
        self.hits = _as_array(res.hits, res.n_hits, HIT_DTYPE)
        self.edit_ops = _as_array(res.edit_ops, res.n_edit_ops, EDIT_OP_DTYPE)
        self.cigar = _as_array(res.cigar, res.n_cigar, np.dtype(np.uint32))
        self.text = _as_array(res.text, res.n_text, np.dtype(np.uint8)).tobytes()
        self.timing = dict(
            h2d=res.ms_h2d, prologue=res.ms_prologue, search=res.ms_search, epilogue=res.ms_epilogue,
            d2h=res.ms_d2h, total=res.ms_total,
        )
        self.gpu_launches = int(res.gpu_launches)
        self.xa = None  # filled by the caller when available

    def __len__(self):
        return len(self.records)

    # -- convenience decoders ---------------------------------------------------------------
    def cigar_str(self, off, n):
        out = []
        for v in self.cigar[off : off + n]:
            out.append("%d%s" % (int(v) >> 4, "MID"[int(v) & 15]))
        return "".join(out)

    def md_str(self, off, n):
        return self.text[off : off + n].decode()

    def record_summary(self, i):
        r = self.records[i]
        if not r["mapped"]:
            return dict(mapped=False, mapq=int(r["mapq"]))
        return dict(
            mapped=True,
            tid=int(r["tid"]),
            pos=int(r["pos"]),
            strand=int(r["strand"]),
            mapq=int(r["mapq"]),
            AS=float(r["alignment_score"]),
            cigar=self.cigar_str(int(r["cigar_off"]), int(r["cigar_len"])),
            md=self.md_str(int(r["md_off"]), int(r["md_len"])),
            nm=int(r["nm"]),
            x0=int(r["x0"]),
            x1=int(r["x1"]),
            xs=float(r["xs"]),
            xt=chr(int(r["xt"])),
        )

    def hits_of(self, i):
        r = self.records[i]
        out = []
        for h in self.hits[int(r["hit_off"]) : int(r["hit_off"]) + int(r["n_hits"])]:
            ops = self.edit_ops[int(h["edit_off"]) : int(h["edit_off"]) + int(h["edit_len"])]
            out.append(
                dict(
                    lower=int(h["lower"]),
                    lower_rev=int(h["lower_rev"]),
                    size=int(h["size"]),
                    score=float(h["alignment_score"]),
                    ops=[(int(o["pos"]), int(o["kind"]), int(o["base"])) for o in ops],
                )
            )
        return out


def pack_reads(seqs, quals):
    """list[bytes], list[bytes|list[int]] -> (seq u8, qual u8, offsets u64) numpy arrays."""
    n = len(seqs)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    for i, s in enumerate(seqs):
        offsets[i + 1] = offsets[i] + len(s)
    seq = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    qual = np.frombuffer(b"".join(bytes(bytearray(q)) for q in quals), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    assert len(seq) == len(qual) == int(offsets[-1])
    return seq, qual, offsets


def results_struct(batch_result):
    """abi.BatchResult (numpy copies) -> (Results struct pointing into them, keepalive list)."""
    r = Results()
    recs = np.ascontiguousarray(batch_result.records)
    hits = np.ascontiguousarray(batch_result.hits)
    ops = np.ascontiguousarray(batch_result.edit_ops)
    cig = np.ascontiguousarray(batch_result.cigar, dtype=np.uint32)
    text = np.frombuffer(batch_result.text, dtype=np.uint8).copy() if len(batch_result.text) else np.zeros(1, np.uint8)
    r.n_reads = len(recs)
    r.records = C.cast(recs.ctypes.data, C.POINTER(Record))
    r.hits = C.cast(hits.ctypes.data, C.POINTER(Hit)); r.n_hits = len(hits)
    r.edit_ops = C.cast(ops.ctypes.data, C.POINTER(EditOp)); r.n_edit_ops = len(ops)
    r.cigar = C.cast(cig.ctypes.data, C.POINTER(C.c_uint32)); r.n_cigar = len(cig)
    r.text = C.cast(text.ctypes.data, C.POINTER(C.c_char)); r.n_text = len(batch_result.text)
    return r, [recs, hits, ops, cig, text]
