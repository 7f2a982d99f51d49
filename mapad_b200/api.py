"""Python binding of the C ABI (include/mapad_gpu.h) in mapad_b200/libmapad_gpu.so.

Mirrors the reference's seam for the hot path: `Index` ~ the loaded index files
(src/index/mod.rs:212-239), `Mapper.map_batch` ~ the per-chunk body of `run_inner`
(src/map/mapping.rs:151-288: k_mismatch_search + intervals_to_bam for every read of a chunk).
There is no CPU fallback: without the built CUDA library or without a CUDA device this module
raises instead of computing anything.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAPAD_GPU_LIB", os.path.join(_HERE, "libmapad_gpu.so"))  # override only for A/B experiments
_lib = None

ERRORS = {0: "OK", -1: "EINVAL", -2: "ENODEV", -3: "ECUDA", -4: "ENOMEM", -5: "EINDEX", -6: "EIO", -7: "ELIMIT"}


class MapadError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("mapad_gpu error %s (%d)%s" % (ERRORS.get(code, "?"), code, (": " + msg) if msg else ""))
        self.code = code


def lib():
    """Loads the CUDA extension.  Fails loudly if it has not been built (python -m mapad_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m mapad_b200.build` (needs nvcc); "
                              "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, f32, i32, u8 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_uint8
        P = C.POINTER
        sig = {
            "mapad_abi_version": (i32, []),
            "mapad_abi_sizeof": (u64, [i32]),
            "mapad_params_from_cli": (i32, [P(abi.Params), C.c_char_p, f32, f32, f32, f32, f32, f32, f32, f32, u8, u8, i32, i32]),
            "mapad_sdm_get": (f32, [P(abi.Params), C.c_size_t, C.c_size_t, u8, u8, u8]),
            "mapad_sdm_representative_mismatch_penalty": (f32, [P(abi.Params)]),
            "mapad_bound_allowed_mismatches": (f32, [P(abi.Params), C.c_size_t]),
            "mapad_index_build": (i32, [u64, P(C.c_char_p), P(C.c_char_p), P(u64), u64, P(vp)]),
            "mapad_index_build_on_device": (i32, [u64, P(C.c_char_p), P(C.c_char_p), P(u64), u64, i32, P(vp)]),
            "mapad_index_build_with_draws": (i32, [u64, P(C.c_char_p), P(C.c_char_p), P(u64), C.c_char_p, u64, P(vp)]),
            "mapad_index_from_view": (i32, [P(abi.IndexView), P(vp)]),
            "mapad_index_get_view": (i32, [vp, P(abi.IndexView)]),
            "mapad_index_free": (None, [vp]),
            "mapad_index_save": (i32, [vp, C.c_char_p]),
            "mapad_index_load": (i32, [C.c_char_p, P(vp)]),
            "mapad_format_xa": (C.c_int64, [vp, P(abi.Results), u64, C.c_char_p, u64]),
            "mapad_gpu_create": (i32, [vp, P(abi.Params), i32, P(vp)]),
            "mapad_gpu_index_meta_size": (u64, []),
            "mapad_gpu_export_index": (i32, [vp, vp, P(vp), P(u64)]),
            "mapad_gpu_copy_index_to": (i32, [vp, vp, u64]),
            "mapad_gpu_create_from_device_blob": (i32, [vp, vp, u64, i32, vp, P(abi.Params), i32, P(vp)]),
            "mapad_gpu_clone_to_device": (i32, [vp, i32, P(vp)]),
            "mapad_gpu_plan_handles": (i32, [i32, i32]),
            "mapad_gpu_set_params": (i32, [vp, P(abi.Params)]),
            "mapad_gpu_map_batch": (i32, [vp, P(abi.Reads), u32, P(abi.Results)]),
            "mapad_gpu_set_stream": (i32, [vp, vp]),
            "mapad_gpu_last_error": (C.c_char_p, [vp]),
            "mapad_gpu_destroy": (None, [vp]),
            "mapad_gpu_gather_peak": (i32, [i32, u64, u32, u64, P(C.c_double)]),
            "mapad_gpu_debug_libm": (i32, [i32, i32, i32, u64, vp, vp]),
            "mapad_input_open": (i32, [C.c_char_p, P(vp)]),
            "mapad_input_is_bam": (i32, [vp]),
            "mapad_input_header_text": (C.c_char_p, [vp]),
            "mapad_input_next_chunk": (i32, [vp, u64, P(vp)]),
            "mapad_input_close": (None, [vp]),
            "mapad_chunk_aux": (i32, [vp, P(vp), P(vp)]),
            "mapad_bam_open_with_header": (i32, [C.c_char_p, vp, C.c_char_p, C.c_char_p, i32, C.c_char_p, P(vp)]),
            "mapad_bam_write_chunk_aux": (i32, [vp, vp, P(abi.Reads), vp, vp, vp, vp, vp, P(abi.Results)]),
            "mapad_chunk_view": (u64, [vp, P(abi.Reads), P(vp), P(vp), P(vp), P(u64)]),
            "mapad_chunk_free": (None, [vp]),
            "mapad_bam_open": (i32, [C.c_char_p, vp, C.c_char_p, C.c_char_p, i32, P(vp)]),
            "mapad_bam_write_chunk": (i32, [vp, vp, P(abi.Reads), vp, vp, vp, P(abi.Results)]),
            "mapad_bam_close": (i32, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "mapad_abi_version", "mapad_abi_sizeof", "mapad_params_from_cli", "mapad_sdm_get", "mapad_sdm_representative_mismatch_penalty",
    "mapad_bound_allowed_mismatches", "mapad_index_build", "mapad_index_build_on_device", "mapad_index_build_with_draws", "mapad_index_from_view",
    "mapad_index_get_view", "mapad_index_free", "mapad_index_save", "mapad_index_load", "mapad_format_xa", "mapad_gpu_create", "mapad_gpu_index_meta_size",
    "mapad_gpu_export_index", "mapad_gpu_copy_index_to", "mapad_gpu_create_from_device_blob", "mapad_gpu_clone_to_device", "mapad_gpu_plan_handles", "mapad_gpu_set_params", "mapad_gpu_map_batch",
    "mapad_gpu_set_stream", "mapad_gpu_last_error", "mapad_gpu_destroy", "mapad_gpu_gather_peak", "mapad_gpu_debug_libm",
    "mapad_chunk_view", "mapad_chunk_free", "mapad_bam_open",
    "mapad_bam_write_chunk", "mapad_bam_close", "mapad_input_open", "mapad_input_is_bam", "mapad_input_header_text", "mapad_input_next_chunk",
    "mapad_input_close", "mapad_chunk_aux", "mapad_bam_open_with_header", "mapad_bam_write_chunk_aux",
]


def _check(rc, handle=None):
    if rc != 0:
        msg = ""
        if handle:
            m = lib().mapad_gpu_last_error(handle)
            msg = m.decode() if m else ""
        raise MapadError(rc, msg)


def params_from_cli(library="single_stranded", p=0.03, f=0.5, t=0.5, d=0.02, s=1.0, D=0.02, i=0.001, x=0.5,
                    gap_dist_ends=5, max_num_gaps_open=2, ignore_base_quality=False, no_search_limit_recovery=False):
    """`mapad map` flags (src/main.rs:120-300) -> AlignmentParameters POD (src/main.rs:418-499)."""
    P = abi.Params()
    _check(lib().mapad_params_from_cli(C.byref(P), library.encode(), p, f, t, d, s, D, i, x, gap_dist_ends, max_num_gaps_open,
                                       int(ignore_base_quality), int(no_search_limit_recovery)))
    return P


def representative_mismatch_penalty(params):
    return float(lib().mapad_sdm_representative_mismatch_penalty(C.byref(params)))


def sdm_get(params, i, L, frm, to, q):
    o = lambda v: ord(v) if isinstance(v, str) else v
    return float(lib().mapad_sdm_get(C.byref(params), i, L, o(frm), o(to), q))


def allowed_mismatches(params, L):
    return float(lib().mapad_bound_allowed_mismatches(C.byref(params), L))


class Index:
    """Host-side index (the arrays of the reference's .tbw/.tle/.toc/.trt/.tsa/.tpi/.tos files)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    @classmethod
    def build(cls, contigs, seed=1234, draws=None, device=None):
        """`mapad index` on in-memory contigs: list[(name, sequence)] (src/index/indexing.rs:29-212).
        device=k: suffix sorting on CUDA device k.  That sorter finishes only texts in which no two suffixes share a 43-symbol
        prefix (the i.i.d. synthetic BASELINE genomes); real genomes (repeats, N runs) fall back to the host SA-IS automatically."""
        n = len(contigs)
        names = (C.c_char_p * n)(*[c[0].encode() if isinstance(c[0], str) else c[0] for c in contigs])
        seqs_b = [c[1].encode() if isinstance(c[1], str) else bytes(c[1]) for c in contigs]
        seqs = (C.c_char_p * n)(*seqs_b)
        lens = (C.c_uint64 * n)(*[len(s) for s in seqs_b])
        out = C.c_void_p()
        if draws is not None:
            d = draws.encode() if isinstance(draws, str) else draws
            _check(lib().mapad_index_build_with_draws(n, names, seqs, lens, d, len(d), C.byref(out)))
        elif device is not None:
            _check(lib().mapad_index_build_on_device(n, names, seqs, lens, seed, int(device), C.byref(out)))
        else:
            _check(lib().mapad_index_build(n, names, seqs, lens, seed, C.byref(out)))
        return cls(out.value)

    def save(self, prefix):
        """Writes `<prefix>.tbw .tle .toc .trt .tsa .tpi .tos` as `mapad index` does (src/index/indexing.rs:111-207)."""
        _check(lib().mapad_index_save(self.h, os.fsencode(prefix)))

    @classmethod
    def load(cls, prefix):
        """Reads the index files next to `prefix` (src/index/mod.rs:212-239)."""
        out = C.c_void_p()
        _check(lib().mapad_index_load(os.fsencode(prefix), C.byref(out)))
        return cls(out.value)

    def __del__(self):
        try:
            if self.h:
                lib().mapad_index_free(self.h)
                self.h = None
        except Exception:
            pass

    def view(self):
        v = abi.IndexView()
        _check(lib().mapad_index_get_view(self.h, C.byref(v)))
        return v

    def arrays(self):
        """Copies of the index arrays as numpy (for cross-checks and for handing the index to the oracle)."""
        v = self.view()
        n = int(v.n)
        out = dict(
            n=n,
            bwt=np.ctypeslib.as_array(v.bwt, shape=(n,)).copy(),
            less=[int(x) for x in v.less],
            sentinel_rows=[int(v.sentinel_rows[0]), int(v.sentinel_rows[1])],
            sa_sample=np.ctypeslib.as_array(v.sa_sample, shape=(int(v.n_sa_samples),)).copy(),
            sa_rate=int(v.sa_rate),
            extra_rows=(np.ctypeslib.as_array(v.extra_rows, shape=(int(v.n_extra_rows) * 2,)).copy().reshape(-1, 2)
                        if v.n_extra_rows else np.zeros((0, 2), np.uint64)),
            contigs=[(v.contig_name[i].decode(), int(v.contig_start[i]), int(v.contig_end[i])) for i in range(int(v.n_contigs))],
            orig_pos=np.ctypeslib.as_array(v.orig_pos, shape=(int(v.n_orig),)).copy() if v.n_orig else np.zeros(0, np.uint64),
            orig_sym=np.ctypeslib.as_array(v.orig_sym, shape=(int(v.n_orig),)).copy() if v.n_orig else np.zeros(0, np.uint8),
        )
        return out

    @property
    def contig_names(self):
        v = self.view()
        return [v.contig_name[i].decode() for i in range(int(v.n_contigs))]


def make_reads(seq, qual, offsets, seeds=None, custom_penalties=None):
    """numpy arrays -> (abi.Reads, keepalive tuple)"""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    qual = np.ascontiguousarray(qual, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    R = abi.Reads()
    R.n_reads = len(offsets) - 1
    R.seq = seq.ctypes.data
    R.qual = qual.ctypes.data
    R.offsets = offsets.ctypes.data
    keep = [seq, qual, offsets]
    if seeds is not None:
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        R.seeds = seeds.ctypes.data
        keep.append(seeds)
    if custom_penalties is not None:
        cp = np.ascontiguousarray(custom_penalties, dtype=np.float32)
        R.custom_penalties = cp.ctypes.data
        keep.append(cp)
    return R, keep


def plan_handles(device, n_handles):
    """Tell the library how many handles are about to be created on `device` (workspace = equal shares of free memory)."""
    _check(lib().mapad_gpu_plan_handles(device, n_handles))


class Mapper:
    """One GPU-resident index + parameters; maps batches of reads (one in flight per handle)."""

    def __init__(self, index, params, device=0, _handle=None):
        self.index = index
        self.params = params
        self.device = device
        if _handle is not None:
            self.h = _handle
        else:
            out = C.c_void_p()
            _check(lib().mapad_gpu_create(index.h, C.byref(params), device, C.byref(out)))
            self.h = out

    def close(self):
        if getattr(self, "h", None):
            lib().mapad_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        _check(lib().mapad_gpu_set_params(self.h, C.byref(params)), self.h)
        self.params = params

    def set_stream(self, cuda_stream_ptr):
        _check(lib().mapad_gpu_set_stream(self.h, cuda_stream_ptr), self.h)

    def export_index(self):
        """(meta bytes, device pointer, nbytes) of the GPU-resident index blob (for the NCCL broadcast)."""
        meta = C.create_string_buffer(int(lib().mapad_gpu_index_meta_size()))
        ptr = C.c_void_p()
        nb = C.c_uint64()
        _check(lib().mapad_gpu_export_index(self.h, meta, C.byref(ptr), C.byref(nb)), self.h)
        return meta.raw, ptr.value, int(nb.value)

    def copy_index_to(self, dst_dev_ptr, nbytes):
        _check(lib().mapad_gpu_copy_index_to(self.h, dst_dev_ptr, nbytes), self.h)

    @classmethod
    def from_device_blob(cls, meta_bytes, dev_ptr, nbytes, index, params, device=0, take_ownership=False):
        out = C.c_void_p()
        meta = C.create_string_buffer(meta_bytes, len(meta_bytes))
        _check(lib().mapad_gpu_create_from_device_blob(meta, dev_ptr, nbytes, int(take_ownership), index.h if index else None,
                                                       C.byref(params), device, C.byref(out)))
        return cls(index, params, device, _handle=out)

    def clone(self, device=None):
        """A second handle.  On this handle's own device it shares the GPU-resident index blob (no copy); on another
        GPU of the box the blob is replicated with one peer-to-peer copy.  Handles are independent otherwise (own
        stream, workspace, one batch in flight each), which lets a driver keep several chunks in flight — straggler
        reads of one chunk overlap with the next — and shard chunks over several GPUs."""
        device = self.device if device is None else device
        out = C.c_void_p()
        _check(lib().mapad_gpu_clone_to_device(self.h, device, C.byref(out)), self.h)
        c = Mapper(self.index, self.params, device, _handle=out)
        if device == self.device:
            c._blob_owner = self  # keep the owner of the shared blob alive
        return c

    def map_raw(self, reads_struct, flags=0):
        res = abi.Results()
        _check(lib().mapad_gpu_map_batch(self.h, C.byref(reads_struct) if reads_struct is not None else None, flags, C.byref(res)), self.h)
        return res

    def map_batch(self, seqs=None, quals=None, seeds=None, want_hits=False, packed=None, custom_penalties=None, with_xa=False):
        """Maps a chunk of reads; returns abi.BatchResult in input order."""
        if packed is None:
            packed = abi.pack_reads(seqs, quals)
        R, keep = make_reads(packed[0], packed[1], packed[2], seeds, custom_penalties)
        flags = abi.BATCH_WANT_HITS if want_hits else 0
        res = self.map_raw(R, flags)
        out = abi.BatchResult(res)
        if with_xa and self.index is not None:
            buf = C.create_string_buffer(1 << 16)
            out.xa = []
            for i in range(len(out)):
                k = lib().mapad_format_xa(self.index.h, C.byref(res), i, buf, 1 << 16)
                if k < 0:
                    raise MapadError(int(k))
                out.xa.append(buf.raw[:k].decode())
        del keep
        return out


def gather_peak(device, table_bytes, bytes_per_access=64, n_accesses=1 << 28):
    out = C.c_double()
    _check(lib().mapad_gpu_gather_peak(device, table_bytes, bytes_per_access, n_accesses, C.byref(out)))
    return float(out.value)


def debug_libm(fn, values, iarg=0, device=0):
    """Device evaluation of the glibc restatements (0 log2f, 1 exp2f, 2 log10f, 3 powi) on a float32 array."""
    x = np.ascontiguousarray(values, dtype=np.float32)
    y = np.zeros_like(x)
    _check(lib().mapad_gpu_debug_libm(device, fn, iarg, len(x), x.ctypes.data, y.ctypes.data))
    return y


class ReadChunks:
    """Iterates a FASTQ / FASTQ.GZ / BAM file (format sniffed) in chunks of `batch_size` reads (`--batch_size`,
    src/main.rs:225-232).  Yields (abi.Reads, names_ptr, name_offsets_ptr, flags_ptr, n_reads, chunk_handle); free with
    .free(handle).  `header_text` is the SAM header of a BAM input (None for FASTQ)."""

    def __init__(self, path, batch_size=250_000):
        self.r = C.c_void_p()
        _check(lib().mapad_input_open(os.fsencode(path), C.byref(self.r)))
        self.batch_size = batch_size
        self.skipped = 0
        self.is_bam = bool(lib().mapad_input_is_bam(self.r))
        t = lib().mapad_input_header_text(self.r)
        self.header_text = t.decode(errors="replace") if t is not None else None

    def __iter__(self):
        return self

    def __next__(self):
        while True:
            ch = C.c_void_p()
            _check(lib().mapad_input_next_chunk(self.r, self.batch_size, C.byref(ch)))
            R = abi.Reads()
            names, noff, flags, skipped = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
            n = lib().mapad_chunk_view(ch, C.byref(R), C.byref(names), C.byref(noff), C.byref(flags), C.byref(skipped))
            self.skipped += int(skipped.value)
            if n:
                return R, names, noff, flags, int(n), ch
            lib().mapad_chunk_free(ch)
            if not skipped.value:  # a chunk of nothing but skipped records is not the end of the input
                raise StopIteration

    def free(self, ch):
        lib().mapad_chunk_free(ch)

    def close(self):
        if self.r:
            lib().mapad_input_close(self.r)
            self.r = None


FastqChunks = ReadChunks


class BamWriter:
    def __init__(self, path, index, command_line="", read_group_id=None, force_overwrite=False, src_header_text=None):
        self.w = C.c_void_p()
        self.index = index
        _check(lib().mapad_bam_open_with_header(os.fsencode(path), index.h, command_line.encode(),
                                                read_group_id.encode() if read_group_id else None, int(force_overwrite),
                                                src_header_text.encode() if src_header_text else None, C.byref(self.w)))

    def write_chunk(self, reads_struct, names, name_offsets, flags, results_struct, chunk=None):
        """chunk: the ReadChunks handle the reads came from — its BAM auxiliary fields are carried over."""
        aux, aoff = C.c_void_p(), C.c_void_p()
        if chunk is not None:
            _check(lib().mapad_chunk_aux(chunk, C.byref(aux), C.byref(aoff)))
        _check(lib().mapad_bam_write_chunk_aux(self.w, self.index.h, C.byref(reads_struct), names, name_offsets, flags, aux, aoff,
                                               C.byref(results_struct)))

    def close(self):
        if self.w:
            _check(lib().mapad_bam_close(self.w))
            self.w = None
