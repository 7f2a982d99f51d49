// common.h — PODs shared by the host code and the CUDA kernels.
#pragma once
#include <cstdint>

namespace mapad {

// ---------------------------------------------------------------------------------------------
// Device index blob.  One contiguous allocation in HBM:
//   [occ blocks][sampled SA][extra rows][X ranges][contig table][original symbols]
//
// occ block, narrow layout (n < 2^31): 32 B = one sector
//     u32 cnt[4]   #A,#C,#G,#T in bwt[0 .. block_start)        (bit 31 of cnt[0]: block holds 'X')
//     u32 code[4]  64 symbols, 2 bits each (A=0,C=1,G=2,T=3; '$' and 'X' are stored as 0)
// occ block, wide layout (n >= 2^31): 64 B
//     u64 cnt[4]   (bit 63 of cnt[0]: block holds 'X')
//     u32 code[8]  128 symbols
// This replaces rust-bio's BWT byte vector + Occ checkpoints every 128 rows
// (/root/reference/src/index/indexing.rs:166,188; src/map/fmd_index.rs:22-25): one block fetch
// yields the ranks of all four bases at a row, so the four extensions of FmdExtIterator
// (fmd_index.rs:162-181) cost two block fetches.
// ---------------------------------------------------------------------------------------------
struct IndexMeta {
  uint64_t n;                  // text length
  uint64_t less[8];
  uint64_t sentinel_rows[2];
  uint32_t wide;               // 0: 32 B blocks / u32 SA samples, 1: 64 B blocks / u64 SA samples
  uint32_t sa_rate;
  uint64_t n_blocks;
  uint64_t n_sa, n_extra, n_xranges, n_contigs, n_orig;
  uint64_t off_occ, off_sa, off_extra, off_xranges, off_contigs, off_orig;  // byte offsets into the blob
  uint64_t total_bytes;
  uint64_t n_x_rows;           // total rows whose BWT symbol is 'X'
};

struct XRange {  // maximal run of rows with bwt == 'X'
  uint64_t start, end;   // [start, end)
  uint64_t before;       // number of X rows in earlier ranges
};

// ---------------------------------------------------------------------------------------------
// Alignment parameters as the kernels consume them (AlignmentParameters, src/map/mod.rs:21-31).
// ---------------------------------------------------------------------------------------------
enum { MODEL_SIMPLE = 0, MODEL_TABLE = 1 };  // device view: built-in aDNA model or host-provided penalties
enum { BOUND_CONTINUOUS = 0, BOUND_DISCRETE = 1, BOUND_TEST = 2 };

struct DevParams {
  int32_t model;            // MODEL_SIMPLE / MODEL_TABLE
  int32_t library;          // 0 single stranded, 1 double stranded
  float overhang5, overhang3;
  float ds_rate, ss_rate, divergence;
  int32_t ignore_q;
  float default_q_prob;     // qual2prob(255)
  int32_t start_mode;       // 0: start = len (SimpleAncientDnaModel), 1: start = len / 2, 2: per-read starts provided
  int32_t bound_kind;
  float repr_mm;
  float cutoff;             // continuous
  float test_threshold;     // test bound
  float test_repr_mm;
  float gap_open, gap_extend;
  int32_t gap_dist_ends;
  int32_t max_num_gaps_open;
  int32_t stack_limit_abort;
  uint32_t stack_limit, edit_tree_limit;
  uint32_t bound_table_len; // entries in the per-length table (discrete: k(L); continuous: L^exponent)
};

// K2 -> K3 hand-over and device-side bump allocators / queues
struct ReadMid {
  uint32_t n_hits, hit_off, frames_popped, flags;
};
struct Cursors {
  uint32_t queue_head;
  uint32_t n_deferred;
  uint32_t hit_cursor;
  uint32_t op_cursor;
  uint32_t cigar_cursor;
  uint32_t text_cursor;
  uint32_t overflow;   // hits / edit-op pool overflow
  uint32_t pad;        // cigar / text pool overflow
};

}  // namespace mapad
