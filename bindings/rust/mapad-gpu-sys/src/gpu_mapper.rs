//! Safe wrapper a mapAD maintainer would put next to `src/map/mapping.rs`: one `GpuMapper` per GPU replaces the body of the
//! chunk loop of `run_inner` (src/map/mapping.rs:151-288) — the `par_iter` over reads calling `k_mismatch_search`
//! (:1012-1383) and `intervals_to_bam` (:402-567).  Reading chunks and writing BAM records stay in mapAD.
//! NOT compiled in this repository's image (no Rust toolchain); kept in step with the header by tests/test_abi.py.
use std::ffi::{CStr, CString};
use std::mem::MaybeUninit;
use std::ptr;

use crate::*;

#[derive(Debug)]
pub struct GpuError {
    pub code: i32,
    pub message: String,
}

fn check(code: i32, handle: *const mapad_gpu) -> Result<(), GpuError> {
    if code == MAPAD_OK {
        return Ok(());
    }
    let message = if handle.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(mapad_gpu_last_error(handle)) }.to_string_lossy().into_owned()
    };
    Err(GpuError { code, message })
}

/// What the wrapper needs from mapAD's `Record` (src/map/record.rs:26-34): bases in original orientation and Phred values.
pub trait ReadLike {
    fn sequence(&self) -> &[u8];
    fn base_qualities(&self) -> &[u8];
}

/// The per-read outcome mapAD's `create_bam_record` (mapping.rs:722-927) consumes.
pub struct MappedRead<'a> {
    pub record: &'a mapad_record,
    pub cigar: &'a [u32],
    pub md: &'a [u8],
}

pub struct GpuMapper {
    handle: *mut mapad_gpu,
    index: *mut mapad_index,
    // packed chunk, reused between calls
    seq: Vec<u8>,
    qual: Vec<u8>,
    offsets: Vec<u64>,
    seeds: Vec<u32>,
}

unsafe impl Send for GpuMapper {}

impl GpuMapper {
    /// `prefix`: path of the reference FASTA; the seven index files written by `mapad index` are loaded from next to it
    /// (load_index_from_path / load_suffix_array_from_path ..., src/index/mod.rs:212-239).
    pub fn from_index_files(prefix: &str, params: &mapad_params, device: i32, chunks_in_flight: i32) -> Result<Self, GpuError> {
        check_layout().map_err(|message| GpuError { code: MAPAD_EINVAL, message })?;
        let c_prefix = CString::new(prefix).map_err(|_| GpuError { code: MAPAD_EINVAL, message: "NUL in path".into() })?;
        let mut index = ptr::null_mut();
        check(unsafe { mapad_index_load(c_prefix.as_ptr(), &mut index) }, ptr::null())?;
        check(unsafe { mapad_gpu_plan_handles(device, chunks_in_flight) }, ptr::null())?;
        let mut handle = ptr::null_mut();
        let rc = unsafe { mapad_gpu_create(index, params, device, &mut handle) };
        if rc != MAPAD_OK {
            unsafe { mapad_index_free(index) };
            return check(rc, ptr::null()).map(|_| unreachable!());
        }
        Ok(GpuMapper { handle, index, seq: Vec::new(), qual: Vec::new(), offsets: Vec::new(), seeds: Vec::new() })
    }

    /// A second handle on `device` sharing (same GPU) or replicating (other GPU, one NVLink peer copy) the resident index:
    /// several chunks in flight per GPU, several GPUs per box — replaces the worker farm of src/distributed for one box.
    pub fn clone_to_device(&self, device: i32) -> Result<Self, GpuError> {
        let mut handle = ptr::null_mut();
        check(unsafe { mapad_gpu_clone_to_device(self.handle, device, &mut handle) }, self.handle)?;
        Ok(GpuMapper { handle, index: ptr::null_mut(), seq: Vec::new(), qual: Vec::new(), offsets: Vec::new(), seeds: Vec::new() })
    }

    /// Maps one chunk; `draw_seed` supplies the per-read value that replaces `rng.next_u32()` of mapping.rs:605.
    /// The closure sees every read's outcome in input order (mapping.rs:288-293) while the result buffers are valid.
    pub fn map_chunk<R: ReadLike>(&mut self, reads: &[R], mut draw_seed: impl FnMut() -> u32,
                                  mut consume: impl FnMut(usize, MappedRead<'_>)) -> Result<(), GpuError> {
        self.seq.clear();
        self.qual.clear();
        self.offsets.clear();
        self.seeds.clear();
        self.offsets.push(0);
        for r in reads {
            self.seq.extend_from_slice(r.sequence());
            self.qual.extend_from_slice(r.base_qualities());
            self.offsets.push(self.seq.len() as u64);
            self.seeds.push(draw_seed());
        }
        let input = mapad_reads {
            n_reads: reads.len() as u64,
            seq: self.seq.as_ptr(),
            qual: self.qual.as_ptr(),
            offsets: self.offsets.as_ptr(),
            seeds: self.seeds.as_ptr(),
            custom_penalties: ptr::null(),
        };
        let mut out = MaybeUninit::<mapad_results>::uninit();
        check(unsafe { mapad_gpu_map_batch(self.handle, &input, 0, out.as_mut_ptr()) }, self.handle)?;
        let out = unsafe { out.assume_init() };
        let records = unsafe { std::slice::from_raw_parts(out.records, out.n_reads as usize) };
        let cigar = unsafe { std::slice::from_raw_parts(out.cigar, out.n_cigar as usize) };
        let text = unsafe { std::slice::from_raw_parts(out.text as *const u8, out.n_text as usize) };
        for (i, record) in records.iter().enumerate() {
            let c = &cigar[record.cigar_off as usize..(record.cigar_off + record.cigar_len) as usize];
            let md = &text[record.md_off as usize..(record.md_off + record.md_len) as usize];
            consume(i, MappedRead { record, cigar: c, md });
        }
        Ok(())
    }
}

impl Drop for GpuMapper {
    fn drop(&mut self) {
        unsafe {
            mapad_gpu_destroy(self.handle);
            if !self.index.is_null() {
                mapad_index_free(self.index);
            }
        }
    }
}
