/* mapad_gpu.h — C ABI of the B200-native mapAD hot path.
 *
 * Drop-in seam.  The reference has no FFI; its seam is the pair of generic
 * functions called per read from the batch loop `run_inner`
 *   (/root/reference/src/map/mapping.rs:151-288):
 *     k_mismatch_search(pattern, quals, params, &RtFmdIndex, ..)   mapping.rs:1012-1383
 *     intervals_to_bam(record, hits, &SA, &id_pos_map, &orig_syms, ..)   mapping.rs:402-567
 * and, as a wire contract, TaskSheet -> ResultSheet
 *   (src/map/input_chunk_reader.rs:247-253, src/distributed/mod.rs:22-26).
 * `mapad_gpu_map_batch` replaces exactly that loop body for one chunk of reads: reads in,
 * hit intervals (the ResultSheet payload) and finished alignment-record fields out, in input order.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success
 * or a negative MAPAD_E* code (never throws, never aborts); caller owns inputs; the library owns
 * outputs, which stay valid until the next map_batch / destroy on the same handle; one in-flight
 * batch per handle.  There is NO CPU fallback: without a CUDA device every mapad_gpu_* call
 * fails with MAPAD_ENODEV.
 */
#ifndef MAPAD_GPU_H
#define MAPAD_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAPAD_ABI_VERSION 1

enum {
  MAPAD_OK = 0,
  MAPAD_EINVAL = -1,   /* bad argument */
  MAPAD_ENODEV = -2,   /* no usable CUDA device */
  MAPAD_ECUDA = -3,    /* CUDA runtime error, see mapad_gpu_last_error */
  MAPAD_ENOMEM = -4,
  MAPAD_EINDEX = -5,   /* inconsistent / unsupported index (Error::InvalidIndex, src/errors.rs) */
  MAPAD_EIO = -6,
  MAPAD_ELIMIT = -7    /* internal capacity exceeded even after the retry lane */
};

/* ---------------------------------------------------------------------------------------------
 * Alignment parameters — mirrors AlignmentParameters (src/map/mod.rs:21-31) with the enum
 * dispatch of SequenceDifferenceModelDispatch (src/map/sequence_difference_models.rs:68-72)
 * and MismatchBoundDispatch (src/map/mismatch_bounds.rs:26-30) flattened into one POD.
 * ------------------------------------------------------------------------------------------- */
enum { MAPAD_MODEL_SIMPLE_ADNA = 0, MAPAD_MODEL_VINDIJA_PWM = 1, MAPAD_MODEL_TEST = 2, MAPAD_MODEL_CUSTOM = 3 };
enum { MAPAD_LIB_SINGLE_STRANDED = 0, MAPAD_LIB_DOUBLE_STRANDED = 1 };
enum { MAPAD_BOUND_CONTINUOUS = 0, MAPAD_BOUND_DISCRETE = 1, MAPAD_BOUND_TEST = 2 };

/* Custom SequenceDifferenceModel (the trait of sequence_difference_models.rs:14-62):
 * `get` must return non-positive log-scores; `find_alignment_start` may be NULL (=> len/2). */
typedef float (*mapad_sdm_get_fn)(void* user, size_t i, size_t read_length, uint8_t from, uint8_t to,
                                  uint8_t base_quality);
typedef int16_t (*mapad_sdm_start_fn)(void* user, size_t pattern_length);

typedef struct mapad_params {
  int32_t model_kind;
  /* SimpleAncientDnaModel::new (sequence_difference_models.rs:279-285) */
  int32_t library;
  float five_prime_overhang;    /* DoubleStranded(x): x goes here */
  float three_prime_overhang;
  float ds_deamination_rate;
  float ss_deamination_rate;
  float divergence;             /* already divided by 3, as in main.rs:452 */
  int32_t ignore_base_quality;
  /* TestDifferenceModel (sequence_difference_models.rs:396-401) */
  float test_deam_score, test_mm_score, test_match_score;
  /* custom model */
  mapad_sdm_get_fn custom_get;
  mapad_sdm_start_fn custom_start;
  void* custom_user;
  /* mismatch bound */
  int32_t bound_kind;
  float poisson_threshold, base_error_rate;      /* Discrete::new   (mismatch_bounds.rs:186-190) */
  float cutoff, exponent;                        /* Continuous::new (mismatch_bounds.rs:102)     */
  float test_threshold, test_representative_mm;  /* TestBound       (mismatch_bounds.rs:264-267) */
  /* 0 => derive with get_representative_mismatch_penalty() (sequence_difference_models.rs:16-31) */
  float representative_mismatch_penalty;
  float penalty_gap_open;
  float penalty_gap_extend;
  uint8_t gap_dist_ends;
  uint8_t max_num_gaps_open;
  uint8_t stack_limit_abort;
  uint8_t reserved0;
  uint32_t stack_limit;      /* 0 => STACK_LIMIT      = 2 000 000 (mapping.rs:53) */
  uint32_t edit_tree_limit;  /* 0 => EDIT_TREE_LIMIT  = 10 000 000 (mapping.rs:54) */
} mapad_params;

/* Fills `p` with the CLI defaults of `mapad map` (src/main.rs:180-300, :418-499) for the given
 * flag values, including the derived representative mismatch penalty and gap penalties. */
int mapad_params_from_cli(mapad_params* p, const char* library, float poisson_prob, float f, float t, float d,
                          float s, float divergence, float indel_rate, float gap_extension_fraction,
                          uint8_t gap_dist_ends, uint8_t max_num_gaps_open, int ignore_base_quality,
                          int no_search_limit_recovery);

/* Host-side evaluation of the scoring traits, bit-identical to what the device uses
 * (SequenceDifferenceModel::get, ::get_representative_mismatch_penalty; Discrete::get). */
float mapad_sdm_get(const mapad_params* p, size_t i, size_t read_length, uint8_t from, uint8_t to, uint8_t q);
float mapad_sdm_representative_mismatch_penalty(const mapad_params* p);
float mapad_bound_allowed_mismatches(const mapad_params* p, size_t read_length);

/* ---------------------------------------------------------------------------------------------
 * Host index — the content of the reference's index files (.tbw .tle .toc .trt .tsa .tpi .tos;
 * src/index/mod.rs:212-239, src/index/indexing.rs:43-212) as in-memory arrays.
 * ------------------------------------------------------------------------------------------- */
typedef struct mapad_index mapad_index;

typedef struct mapad_index_view {
  uint64_t n;                 /* text length = 2*G + 2 */
  const uint8_t* bwt;         /* n rank bytes ($=0 A=1 C=2 G=3 T=4 X=5)          (.tbw) */
  uint64_t less[8];           /* less[c] = #symbols < c                          (.tle) */
  uint64_t sentinel_rows[2];  /* RtFmdIndex::sentinel_occ (fmd_index.rs:38-47)          */
  const uint64_t* sa_sample;  /* SA[i*rate]                                      (.tsa) */
  uint64_t n_sa_samples;
  uint64_t sa_rate;
  const uint64_t* extra_rows; /* (row, text position) pairs, sorted by row       (.tsa) */
  uint64_t n_extra_rows;
  uint64_t n_contigs;         /* FastaIdPositions                                (.tpi) */
  const uint64_t* contig_start;
  const uint64_t* contig_end; /* inclusive */
  const char* const* contig_name;
  const uint64_t* orig_pos;   /* OriginalSymbols: sorted forward-strand positions (.tos) */
  const uint8_t* orig_sym;
  uint64_t n_orig;
} mapad_index_view;

/* `mapad index` (indexing.rs:29-212): upper-case, IUPAC runs >= 20 -> 'X', shorter ones replaced by
 * a random compatible base drawn from a generator seeded with `seed`, text = fwd $ revcomp $,
 * suffix array, BWT, SA sampled every 32 rows.  `sequences[i]` need not be NUL terminated. */
int mapad_index_build(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                      const uint64_t* lengths, uint64_t seed, mapad_index** out);
/* Same result, but the suffix sorting runs on CUDA device `device` (prefix-key radix sort; falls back to the host
 * SA-IS for texts in which two suffixes share a 43-symbol prefix, i.e. for every real genome: repeats, N runs >= 20 bp).
 * It is what makes the hg19-scale SYNTHETIC (i.i.d.) BASELINE reference indexable in a minute. */
int mapad_index_build_on_device(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                                const uint64_t* lengths, uint64_t seed, int device, mapad_index** out);
/* Same, but ambiguous symbols in short runs are replaced by the bytes of `replacement_draws`
 * (consumed in text order) — lets tests reproduce a given `mapad index` outcome. */
int mapad_index_build_with_draws(uint64_t n_contigs, const char* const* names, const char* const* sequences,
                                 const uint64_t* lengths, const char* replacement_draws, uint64_t n_draws,
                                 mapad_index** out);
/* Wrap arrays produced elsewhere (e.g. a device-side builder). Copies everything. */
int mapad_index_from_view(const mapad_index_view* v, mapad_index** out);
int mapad_index_get_view(const mapad_index* ix, mapad_index_view* out);
void mapad_index_free(mapad_index* ix);
/* The reference's on-disk index: `<prefix>.tbw .tle .toc .trt .tsa .tpi .tos`, each a Snappy frame stream around a
 * bincode `Item{version: u8 = 5, data}` (writers: src/index/indexing.rs:111-207; readers: src/index/mod.rs:212-239,
 * src/index/versioned_index.rs:46-56).  `save` writes all seven (uncompressed frame chunks); `load` reads .tbw .tsa
 * .tpi .tos, validates .tle/.trt against them (ranks are re-derived from the BWT, so .toc is not read) and returns
 * MAPAD_EINDEX on a version mismatch (Error::IndexVersionMismatch) or inconsistent content, MAPAD_EIO on I/O errors. */
int mapad_index_save(const mapad_index* ix, const char* prefix);
int mapad_index_load(const char* prefix, mapad_index** out);

/* ---------------------------------------------------------------------------------------------
 * Batch in / out
 * ------------------------------------------------------------------------------------------- */
typedef struct mapad_reads {
  uint64_t n_reads;
  const uint8_t* seq;        /* ASCII bases, original read orientation, concatenated (Record::sequence) */
  const uint8_t* qual;       /* Phred values (not +33), concatenated       (Record::base_qualities) */
  const uint64_t* offsets;   /* n_reads + 1 entries                                                  */
  const uint32_t* seeds;     /* per-read seed replacing `rng.next_u32()` (mapping.rs:605); NULL => 0 */
  /* MAPAD_MODEL_CUSTOM only: optional precomputed penalties get(i,len,b,read[i],q[i]) for b = A,C,G,T,
   * 4 floats per base, concatenated like seq.  NULL => the callback in mapad_params is evaluated. */
  const float* custom_penalties;
} mapad_reads;

/* EditOperation (src/map/record.rs:226-231) */
enum { MAPAD_ED_INSERTION = 0, MAPAD_ED_DELETION = 1, MAPAD_ED_MATCH = 2, MAPAD_ED_MISMATCH = 3 };
typedef struct mapad_edit_op {
  uint16_t pos;   /* read position carried by the operation */
  uint8_t kind;
  uint8_t base;   /* reference base (ASCII) for deletion / mismatch, else 0 */
} mapad_edit_op;

/* HitInterval (src/map/mod.rs:34-39) */
typedef struct mapad_hit {
  uint64_t lower, lower_rev, size;  /* RtBiInterval (fmd_index.rs:185-189) */
  float alignment_score;
  uint32_t edit_off;                /* span in mapad_results::edit_ops (track order of record.rs:465-500) */
  uint32_t edit_len;
  uint32_t reserved;
} mapad_hit;

/* One alternative position reported through XA (mapping.rs:436-491) */
typedef struct mapad_alt {
  int32_t tid;
  int32_t strand;        /* 0 '+', 1 '-' */
  int64_t pos;           /* 0-based position on the contig */
  uint32_t cigar_off, cigar_len;
  uint32_t md_off, md_len;
  int32_t nm;
  float alignment_score;
  uint64_t interval_size;
} mapad_alt;

/* What intervals_to_bam decides for one read (mapping.rs:402-567, :658-718; record.rs:282-428) */
typedef struct mapad_record {
  int32_t mapped;        /* 0 => unmapped (flag 4, MAPQ 0) */
  int32_t tid;
  int64_t pos;           /* 0-based leftmost position on the contig (BAM POS) */
  int32_t strand;        /* 0 forward, 1 reverse (flag 16) */
  int32_t mapq;
  float alignment_score; /* AS:f */
  int32_t nm;            /* NM:i */
  int32_t x0, x1;
  float xs;              /* XS:f, meaningful when x1 > 0 */
  int32_t xt;            /* 'U' / 'R' / 'N' */
  uint32_t cigar_off, cigar_len;   /* span in mapad_results::cigar, BAM encoding (len << 4 | op; M=0 I=1 D=2) */
  uint32_t md_off, md_len;         /* span in mapad_results::text */
  uint32_t n_alts;                 /* 0..2 */
  mapad_alt alts[2];
  uint32_t hit_off, n_hits;        /* span in mapad_results::hits (BinaryHeap backing-vector order) */
  uint64_t best_lower, best_lower_rev, best_size;   /* interval of the reported hit */
  uint64_t absolute_pos;           /* forward-strand text position */
  /* work counters */
  uint32_t frames_popped;          /* P */
  uint32_t d_ext_steps;            /* E */
  uint32_t lf_steps;               /* W */
  uint32_t flags;                  /* bit0: stack/tree limit was hit; bit1: went through the retry lane */
} mapad_record;

typedef struct mapad_results {
  uint64_t n_reads;
  const mapad_record* records;
  const mapad_hit* hits;       uint64_t n_hits;
  const mapad_edit_op* edit_ops; uint64_t n_edit_ops;
  const uint32_t* cigar;       uint64_t n_cigar;
  const char* text;            uint64_t n_text;
  /* timing of the last batch, milliseconds of device time (CUDA events on the library's stream) */
  float ms_h2d, ms_prologue, ms_search, ms_epilogue, ms_d2h, ms_total;
  uint64_t gpu_launches;       /* kernels launched for this batch */
} mapad_results;

/* XA:Z string of a record, formatted like mapping.rs:475-488 ("chr,+pos,CIGAR,MD,NM,size,AS;").
 * Returns the number of bytes written (excluding the NUL), or a negative error code. */
int64_t mapad_format_xa(const mapad_index* ix, const mapad_results* res, uint64_t read_idx, char* buf, uint64_t cap);

enum {
  MAPAD_BATCH_WANT_HITS = 1u,      /* also return every hit interval with its edit operations */
  MAPAD_BATCH_RESIDENT = 2u,       /* `in` is ignored: re-run on the batch the previous call left resident in HBM */
  MAPAD_BATCH_NO_D2H = 4u,         /* leave results on the device (bench: kernel-only timing) */
  MAPAD_BATCH_UPLOAD_ONLY = 8u     /* stage `in` in HBM and return; run it later with MAPAD_BATCH_RESIDENT */
};

typedef struct mapad_gpu mapad_gpu;

/* Re-lays the index out for the GPU (interleaved count + 2-bit BWT blocks, packed sampled SA,
 * contig / original-symbol tables) and uploads it to `device`. */
int mapad_gpu_create(const mapad_index* ix, const mapad_params* params, int device, mapad_gpu** out);
/* Multi-GPU start-up: rank 0 exports the device-resident index blob, the launcher broadcasts it
 * (NCCL over NVLink) into a buffer on every other GPU, and those ranks adopt it without touching
 * the host index.  `meta` is an opaque POD of mapad_gpu_index_meta_size() bytes. */
uint64_t mapad_gpu_index_meta_size(void);
int mapad_gpu_export_index(mapad_gpu* h, void* meta_out, void** dev_ptr_out, uint64_t* dev_bytes_out);
/* Device-to-device copy of the index blob into a caller-owned device buffer (e.g. a torch tensor that
 * torch.distributed then broadcasts). */
int mapad_gpu_copy_index_to(mapad_gpu* h, void* dst_dev_ptr, uint64_t dst_bytes);
int mapad_gpu_create_from_device_blob(const void* meta, void* dev_ptr, uint64_t dev_bytes, int take_ownership,
                                      const mapad_index* contigs_and_symbols, const mapad_params* params,
                                      int device, mapad_gpu** out);
/* Single-process multi-GPU: a second handle for `device`.  On the source handle's own GPU the resident index blob is
 * shared; on another GPU of the box it is replicated with ONE peer-to-peer copy (NVLink) — the in-process counterpart of
 * the NCCL broadcast above.  Replaces the per-worker index loading of the reference's worker farm
 * (/root/reference/src/distributed/worker.rs:45-215) for single-box runs; the caller shards each chunk of reads over the
 * handles and merges the results in input order like the dispatcher (src/distributed/dispatcher.rs:341-379). */
int mapad_gpu_clone_to_device(mapad_gpu* src, int device, mapad_gpu** out);
/* Announces how many handles the caller is about to create on `device`.  All handles of a device share ONE search
 * workspace (the chunk pool the per-read heaps / edit trees grow in — the thread-local scratch of mapping.rs:146-149),
 * allocated with the first handle: 75 % of the free device memory, but leaving 1 GiB per announced handle for their batch
 * buffers.  MAPAD_WS_BYTES overrides the size. */
int mapad_gpu_plan_handles(int device, int n_handles);
int mapad_gpu_set_params(mapad_gpu* h, const mapad_params* params);
int mapad_gpu_map_batch(mapad_gpu* h, const mapad_reads* in, uint32_t flags, mapad_results* out);
/* Uses the caller's CUDA stream (a cudaStream_t cast to void*) for all subsequent work; NULL = own stream. */
int mapad_gpu_set_stream(mapad_gpu* h, void* cuda_stream);
const char* mapad_gpu_last_error(const mapad_gpu* h);
void mapad_gpu_destroy(mapad_gpu* h);

/* Roofline denominator: independent uniformly random `bytes_per_access`-byte loads
 * over a table of `table_bytes`; returns achieved GB/s. */
int mapad_gpu_gather_peak(int device, uint64_t table_bytes, uint32_t bytes_per_access, uint64_t n_accesses,
                          double* gbps_out);

/* ---------------------------------------------------------------------------------------------
 * Callers and data formats either side of the hot path: FASTQ(.GZ) / BAM in, BAM out.
 * ------------------------------------------------------------------------------------------- */
/* Input reader.  The format (FASTQ, FASTQ.GZ, BAM) is sniffed from the decompressed first bytes as in
 * src/map/input_chunk_reader.rs:42-172 (CRAM: MAPAD_EINVAL).  Record normalisation of src/map/record.rs:138-215:
 * FASTQ is upper-cased, Phred+33 removed, flags 0; BAM reads stored reverse-complemented (flag 16) are turned back,
 * flags and auxiliary fields are kept.  Chunking and bad-record skipping of input_chunk_reader.rs:176-244. */
int mapad_input_open(const char* path, void** reader_out);
int mapad_input_is_bam(void* reader);
const char* mapad_input_header_text(void* reader); /* SAM header text of a BAM input, NULL otherwise */
int mapad_input_next_chunk(void* reader, uint64_t max_reads, void** chunk_out);
void mapad_input_close(void* reader);
uint64_t mapad_chunk_view(void* chunk, mapad_reads* reads, const char** names, const uint64_t** name_offsets,
                          const uint16_t** flags, uint64_t* skipped);
/* raw BAM auxiliary fields of the chunk's reads: read i owns aux[aux_offsets[i] .. aux_offsets[i+1]) */
int mapad_chunk_aux(void* chunk, const uint8_t** aux, const uint64_t** aux_offsets);
void mapad_chunk_free(void* chunk);
/* BAM writer: create_bam_header (src/map/mapping.rs:300-398; `src_header_text` = header of a BAM input, its @PG
 * chain, @RG and @CO lines are carried over) and create_bam_record (:722-927; `aux` = the input's auxiliary fields,
 * copied except for the tag filter of :829-846) from the per-read fields of a mapad_results; records are written in
 * input order (mapping.rs:291-293). */
int mapad_bam_open(const char* path, const mapad_index* index, const char* command_line, const char* read_group_id,
                   int force_overwrite, void** writer_out);
int mapad_bam_open_with_header(const char* path, const mapad_index* index, const char* command_line,
                               const char* read_group_id, int force_overwrite, const char* src_header_text,
                               void** writer_out);
int mapad_bam_write_chunk(void* writer, const mapad_index* index, const mapad_reads* reads, const char* names,
                          const uint64_t* name_offsets, const uint16_t* in_flags, const mapad_results* res);
int mapad_bam_write_chunk_aux(void* writer, const mapad_index* index, const mapad_reads* reads, const char* names,
                              const uint64_t* name_offsets, const uint16_t* in_flags, const uint8_t* aux,
                              const uint64_t* aux_offsets, const mapad_results* res);
int mapad_bam_close(void* writer);

/* Test hook: evaluates the device restatements of the glibc float functions the reference reaches through
 * f32::{log2, exp2, log10} (fn = 0, 1, 2) and compiler-rt's powi (fn = 3, exponent in `iarg`) on `n` host values. */
int mapad_gpu_debug_libm(int device, int fn, int iarg, uint64_t n, const float* in, float* out);

int mapad_abi_version(void);
/* sizeof of the ABI PODs, so that bindings can verify their mirrors: 0 params, 1 reads, 2 edit_op, 3 hit, 4 alt,
 * 5 record, 6 results, 7 index_view */
uint64_t mapad_abi_sizeof(int what);

#ifdef __cplusplus
}
#endif
#endif /* MAPAD_GPU_H */
