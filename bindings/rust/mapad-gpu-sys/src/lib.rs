//! Raw bindings to `include/mapad_gpu.h` (ABI version 1), field for field and symbol for symbol.
//!
//! NOT compiled in this repository's image (there is no Rust toolchain); `tests/test_abi.py::test_rust_binding_covers_header`
//! keeps the symbol list and the POD field counts in step with the header, and `mapad_abi_sizeof` lets the crate verify its
//! mirrors at start-up (`check_layout`).  The seam these functions replace in mapAD is the per-chunk loop body of
//! `run_inner` (src/map/mapping.rs:151-288); see INTEGRATION.md §1-2 and `gpu_mapper.rs` next to this file.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

pub mod gpu_mapper;

pub const MAPAD_ABI_VERSION: c_int = 1;

pub const MAPAD_OK: c_int = 0;
pub const MAPAD_EINVAL: c_int = -1;
pub const MAPAD_ENODEV: c_int = -2;
pub const MAPAD_ECUDA: c_int = -3;
pub const MAPAD_ENOMEM: c_int = -4;
pub const MAPAD_EINDEX: c_int = -5;
pub const MAPAD_EIO: c_int = -6;
pub const MAPAD_ELIMIT: c_int = -7;

pub const MAPAD_MODEL_SIMPLE_ADNA: i32 = 0;
pub const MAPAD_MODEL_VINDIJA_PWM: i32 = 1;
pub const MAPAD_MODEL_TEST: i32 = 2;
pub const MAPAD_MODEL_CUSTOM: i32 = 3;
pub const MAPAD_LIB_SINGLE_STRANDED: i32 = 0;
pub const MAPAD_LIB_DOUBLE_STRANDED: i32 = 1;
pub const MAPAD_BOUND_CONTINUOUS: i32 = 0;
pub const MAPAD_BOUND_DISCRETE: i32 = 1;
pub const MAPAD_BOUND_TEST: i32 = 2;

pub const MAPAD_ED_INSERTION: u8 = 0;
pub const MAPAD_ED_DELETION: u8 = 1;
pub const MAPAD_ED_MATCH: u8 = 2;
pub const MAPAD_ED_MISMATCH: u8 = 3;

pub const MAPAD_BATCH_WANT_HITS: u32 = 1;
pub const MAPAD_BATCH_RESIDENT: u32 = 2;
pub const MAPAD_BATCH_NO_D2H: u32 = 4;
pub const MAPAD_BATCH_UPLOAD_ONLY: u32 = 8;

pub type mapad_sdm_get_fn =
    Option<unsafe extern "C" fn(user: *mut c_void, i: usize, read_length: usize, from: u8, to: u8, base_quality: u8) -> c_float>;
pub type mapad_sdm_start_fn = Option<unsafe extern "C" fn(user: *mut c_void, pattern_length: usize) -> i16>;

/// AlignmentParameters (src/map/mod.rs:21-31) with both enum dispatches flattened.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct mapad_params {
    pub model_kind: i32,
    pub library: i32,
    pub five_prime_overhang: c_float,
    pub three_prime_overhang: c_float,
    pub ds_deamination_rate: c_float,
    pub ss_deamination_rate: c_float,
    pub divergence: c_float,
    pub ignore_base_quality: i32,
    pub test_deam_score: c_float,
    pub test_mm_score: c_float,
    pub test_match_score: c_float,
    pub custom_get: mapad_sdm_get_fn,
    pub custom_start: mapad_sdm_start_fn,
    pub custom_user: *mut c_void,
    pub bound_kind: i32,
    pub poisson_threshold: c_float,
    pub base_error_rate: c_float,
    pub cutoff: c_float,
    pub exponent: c_float,
    pub test_threshold: c_float,
    pub test_representative_mm: c_float,
    pub representative_mismatch_penalty: c_float,
    pub penalty_gap_open: c_float,
    pub penalty_gap_extend: c_float,
    pub gap_dist_ends: u8,
    pub max_num_gaps_open: u8,
    pub stack_limit_abort: u8,
    pub reserved0: u8,
    pub stack_limit: u32,
    pub edit_tree_limit: u32,
}

#[repr(C)]
pub struct mapad_index {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct mapad_gpu {
    _opaque: [u8; 0],
}

/// The arrays of the reference's index files (.tbw .tle .toc .trt .tsa .tpi .tos).
#[repr(C)]
pub struct mapad_index_view {
    pub n: u64,
    pub bwt: *const u8,
    pub less: [u64; 8],
    pub sentinel_rows: [u64; 2],
    pub sa_sample: *const u64,
    pub n_sa_samples: u64,
    pub sa_rate: u64,
    pub extra_rows: *const u64,
    pub n_extra_rows: u64,
    pub n_contigs: u64,
    pub contig_start: *const u64,
    pub contig_end: *const u64,
    pub contig_name: *const *const c_char,
    pub orig_pos: *const u64,
    pub orig_sym: *const u8,
    pub n_orig: u64,
}

/// One chunk of reads (TaskSheet payload, src/map/input_chunk_reader.rs:247-253).
#[repr(C)]
pub struct mapad_reads {
    pub n_reads: u64,
    pub seq: *const u8,
    pub qual: *const u8,
    pub offsets: *const u64,
    pub seeds: *const u32,
    pub custom_penalties: *const c_float,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mapad_edit_op {
    pub pos: u16,
    pub kind: u8,
    pub base: u8,
}

/// HitInterval (src/map/mod.rs:34-39).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct mapad_hit {
    pub lower: u64,
    pub lower_rev: u64,
    pub size: u64,
    pub alignment_score: c_float,
    pub edit_off: u32,
    pub edit_len: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mapad_alt {
    pub tid: i32,
    pub strand: i32,
    pub pos: i64,
    pub cigar_off: u32,
    pub cigar_len: u32,
    pub md_off: u32,
    pub md_len: u32,
    pub nm: i32,
    pub alignment_score: c_float,
    pub interval_size: u64,
}

/// What intervals_to_bam decides for one read (src/map/mapping.rs:402-567).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct mapad_record {
    pub mapped: i32,
    pub tid: i32,
    pub pos: i64,
    pub strand: i32,
    pub mapq: i32,
    pub alignment_score: c_float,
    pub nm: i32,
    pub x0: i32,
    pub x1: i32,
    pub xs: c_float,
    pub xt: i32,
    pub cigar_off: u32,
    pub cigar_len: u32,
    pub md_off: u32,
    pub md_len: u32,
    pub n_alts: u32,
    pub alts: [mapad_alt; 2],
    pub hit_off: u32,
    pub n_hits: u32,
    pub best_lower: u64,
    pub best_lower_rev: u64,
    pub best_size: u64,
    pub absolute_pos: u64,
    pub frames_popped: u32,
    pub d_ext_steps: u32,
    pub lf_steps: u32,
    pub flags: u32,
}

/// ResultSheet payload + finished record fields of one chunk; valid until the handle's next call.
#[repr(C)]
pub struct mapad_results {
    pub n_reads: u64,
    pub records: *const mapad_record,
    pub hits: *const mapad_hit,
    pub n_hits: u64,
    pub edit_ops: *const mapad_edit_op,
    pub n_edit_ops: u64,
    pub cigar: *const u32,
    pub n_cigar: u64,
    pub text: *const c_char,
    pub n_text: u64,
    pub ms_h2d: c_float,
    pub ms_prologue: c_float,
    pub ms_search: c_float,
    pub ms_epilogue: c_float,
    pub ms_d2h: c_float,
    pub ms_total: c_float,
    pub gpu_launches: u64,
}

extern "C" {
    pub fn mapad_params_from_cli(p: *mut mapad_params, library: *const c_char, poisson_prob: c_float, f: c_float, t: c_float,
                                 d: c_float, s: c_float, divergence: c_float, indel_rate: c_float, gap_extension_fraction: c_float,
                                 gap_dist_ends: u8, max_num_gaps_open: u8, ignore_base_quality: c_int,
                                 no_search_limit_recovery: c_int) -> c_int;
    pub fn mapad_sdm_get(p: *const mapad_params, i: usize, read_length: usize, from: u8, to: u8, q: u8) -> c_float;
    pub fn mapad_sdm_representative_mismatch_penalty(p: *const mapad_params) -> c_float;
    pub fn mapad_bound_allowed_mismatches(p: *const mapad_params, read_length: usize) -> c_float;

    pub fn mapad_index_build(n_contigs: u64, names: *const *const c_char, sequences: *const *const c_char, lengths: *const u64,
                             seed: u64, out: *mut *mut mapad_index) -> c_int;
    pub fn mapad_index_build_on_device(n_contigs: u64, names: *const *const c_char, sequences: *const *const c_char,
                                       lengths: *const u64, seed: u64, device: c_int, out: *mut *mut mapad_index) -> c_int;
    pub fn mapad_index_build_with_draws(n_contigs: u64, names: *const *const c_char, sequences: *const *const c_char,
                                        lengths: *const u64, replacement_draws: *const c_char, n_draws: u64,
                                        out: *mut *mut mapad_index) -> c_int;
    pub fn mapad_index_from_view(v: *const mapad_index_view, out: *mut *mut mapad_index) -> c_int;
    pub fn mapad_index_get_view(ix: *const mapad_index, out: *mut mapad_index_view) -> c_int;
    pub fn mapad_index_free(ix: *mut mapad_index);
    pub fn mapad_index_save(ix: *const mapad_index, prefix: *const c_char) -> c_int;
    pub fn mapad_index_load(prefix: *const c_char, out: *mut *mut mapad_index) -> c_int;

    pub fn mapad_format_xa(ix: *const mapad_index, res: *const mapad_results, read_idx: u64, buf: *mut c_char, cap: u64) -> i64;

    pub fn mapad_gpu_create(ix: *const mapad_index, params: *const mapad_params, device: c_int, out: *mut *mut mapad_gpu) -> c_int;
    pub fn mapad_gpu_index_meta_size() -> u64;
    pub fn mapad_gpu_export_index(h: *mut mapad_gpu, meta_out: *mut c_void, dev_ptr_out: *mut *mut c_void,
                                  dev_bytes_out: *mut u64) -> c_int;
    pub fn mapad_gpu_copy_index_to(h: *mut mapad_gpu, dst_dev_ptr: *mut c_void, dst_bytes: u64) -> c_int;
    pub fn mapad_gpu_create_from_device_blob(meta: *const c_void, dev_ptr: *mut c_void, dev_bytes: u64, take_ownership: c_int,
                                             contigs_and_symbols: *const mapad_index, params: *const mapad_params, device: c_int,
                                             out: *mut *mut mapad_gpu) -> c_int;
    pub fn mapad_gpu_clone_to_device(src: *mut mapad_gpu, device: c_int, out: *mut *mut mapad_gpu) -> c_int;
    pub fn mapad_gpu_plan_handles(device: c_int, n_handles: c_int) -> c_int;
    pub fn mapad_gpu_set_params(h: *mut mapad_gpu, params: *const mapad_params) -> c_int;
    pub fn mapad_gpu_map_batch(h: *mut mapad_gpu, input: *const mapad_reads, flags: u32, out: *mut mapad_results) -> c_int;
    pub fn mapad_gpu_set_stream(h: *mut mapad_gpu, cuda_stream: *mut c_void) -> c_int;
    pub fn mapad_gpu_last_error(h: *const mapad_gpu) -> *const c_char;
    pub fn mapad_gpu_destroy(h: *mut mapad_gpu);
    pub fn mapad_gpu_gather_peak(device: c_int, table_bytes: u64, bytes_per_access: u32, n_accesses: u64,
                                 gbps_out: *mut c_double) -> c_int;

    pub fn mapad_input_open(path: *const c_char, reader_out: *mut *mut c_void) -> c_int;
    pub fn mapad_input_is_bam(reader: *mut c_void) -> c_int;
    pub fn mapad_input_header_text(reader: *mut c_void) -> *const c_char;
    pub fn mapad_input_next_chunk(reader: *mut c_void, max_reads: u64, chunk_out: *mut *mut c_void) -> c_int;
    pub fn mapad_input_close(reader: *mut c_void);
    pub fn mapad_chunk_view(chunk: *mut c_void, reads: *mut mapad_reads, names: *mut *const c_char,
                            name_offsets: *mut *const u64, flags: *mut *const u16, skipped: *mut u64) -> u64;
    pub fn mapad_chunk_aux(chunk: *mut c_void, aux: *mut *const u8, aux_offsets: *mut *const u64) -> c_int;
    pub fn mapad_chunk_free(chunk: *mut c_void);
    pub fn mapad_bam_open(path: *const c_char, index: *const mapad_index, command_line: *const c_char,
                          read_group_id: *const c_char, force_overwrite: c_int, writer_out: *mut *mut c_void) -> c_int;
    pub fn mapad_bam_open_with_header(path: *const c_char, index: *const mapad_index, command_line: *const c_char,
                                      read_group_id: *const c_char, force_overwrite: c_int, src_header_text: *const c_char,
                                      writer_out: *mut *mut c_void) -> c_int;
    pub fn mapad_bam_write_chunk(writer: *mut c_void, index: *const mapad_index, reads: *const mapad_reads, names: *const c_char,
                                 name_offsets: *const u64, in_flags: *const u16, res: *const mapad_results) -> c_int;
    pub fn mapad_bam_write_chunk_aux(writer: *mut c_void, index: *const mapad_index, reads: *const mapad_reads,
                                     names: *const c_char, name_offsets: *const u64, in_flags: *const u16, aux: *const u8,
                                     aux_offsets: *const u64, res: *const mapad_results) -> c_int;
    pub fn mapad_bam_close(writer: *mut c_void) -> c_int;

    pub fn mapad_gpu_debug_libm(device: c_int, func: c_int, iarg: c_int, n: u64, input: *const c_float, out: *mut c_float) -> c_int;
    pub fn mapad_abi_version() -> c_int;
    pub fn mapad_abi_sizeof(what: c_int) -> u64;
}

/// Verifies the mirrors above against the library (call once at start-up).
pub fn check_layout() -> Result<(), String> {
    use std::mem::size_of;
    if unsafe { mapad_abi_version() } != MAPAD_ABI_VERSION {
        return Err("libmapad_gpu.so: ABI version mismatch".into());
    }
    let want = [
        size_of::<mapad_params>(), size_of::<mapad_reads>(), size_of::<mapad_edit_op>(), size_of::<mapad_hit>(),
        size_of::<mapad_alt>(), size_of::<mapad_record>(), size_of::<mapad_results>(), size_of::<mapad_index_view>(),
    ];
    for (what, w) in want.iter().enumerate() {
        let got = unsafe { mapad_abi_sizeof(what as c_int) } as usize;
        if got != *w {
            return Err(format!("POD {what}: library says {got} bytes, binding has {w}"));
        }
    }
    Ok(())
}
