// host_io.cpp — callers and data formats on either side of the hot path:
//   * FASTQ / FASTQ.GZ / BAM reader (format sniffed like /root/reference/src/map/input_chunk_reader.rs:42-172) with the
//     reference's Record normalisation (src/map/record.rs:138-215: FASTQ upper-cased, Phred+33 removed, flags 0; BAM
//     reads flagged reverse-complemented are turned back, flags and auxiliary fields are kept).  CRAM is not read.
//   * BAM writer: create_bam_header (src/map/mapping.rs:300-398) and create_bam_record (:722-927) on top of the
//     per-read fields the device produced, BGZF-compressed with zlib.
// The reference uses the noodles crate for both; only the byte formats (SAM/BAM spec) are shared.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "host_index.hpp"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// FASTQ
// ---------------------------------------------------------------------------------------------------------------
struct ReadChunk {
  std::vector<uint8_t> seq, qual;
  std::vector<uint64_t> offsets{0};
  std::vector<char> names;
  std::vector<uint64_t> name_offsets{0};
  std::vector<uint16_t> flags;
  std::vector<uint8_t> aux;               // raw BAM auxiliary fields of every read (empty for FASTQ input)
  std::vector<uint64_t> aux_offsets{0};
  uint64_t skipped = 0;
};

struct InputReader {  // the sniffing reader for FASTQ, FASTQ.GZ and BAM
  gzFile f = nullptr;  // gzopen reads plain files transparently, and BGZF is a series of gzip members
  bool is_bam = false;
  std::string header_text;  // BAM input: the SAM header text
  std::string prefix;       // bytes consumed while sniffing a FASTQ file
  std::string line;
  bool getline() {
    line.clear();
    if (!prefix.empty()) {
      size_t nl = prefix.find('\n');
      if (nl != std::string::npos) {
        line = prefix.substr(0, nl);
        prefix.erase(0, nl + 1);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
      }
      line = prefix;
      prefix.clear();
    }
    char buf[65536];
    while (true) {
      if (!gzgets(f, buf, sizeof buf)) return !line.empty();
      line += buf;
      if (!line.empty() && line.back() == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); return true; }
      if (gzeof(f)) return true;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// BGZF / BAM
// ---------------------------------------------------------------------------------------------------------------
// One BGZF block (SAM spec 4.1) from up to 0xff00 input bytes; `out` must hold 65536 + 1024 bytes.  Returns its size, 0 on error.
size_t bgzf_compress(const uint8_t* data, size_t n, uint8_t* out, int level) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return 0;
  zs.next_in = const_cast<uint8_t*>(data);
  zs.avail_in = (uInt)n;
  zs.next_out = out + 18;
  zs.avail_out = 65536 + 1024 - 18 - 8;
  const int rc = deflate(&zs, Z_FINISH);
  const size_t clen = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) return 0;
  const size_t bsize = clen + 18 + 8;
  if (bsize > 65536) return 0;
  static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
  memcpy(out, hdr, 16);
  out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
  const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n);
  const uint32_t isize = (uint32_t)n;
  memcpy(out + 18 + clen, &crc, 4);
  memcpy(out + 18 + clen + 4, &isize, 4);
  return bsize;
}

struct BamWriter {
  static constexpr size_t BLOCK = 0xff00;
  FILE* f = nullptr;
  std::vector<uint8_t> pending;  // uncompressed bytes not yet written (less than one block between calls)
  std::vector<std::string> contig_names;
  std::string read_group;  // ID or empty
  int level = 6;
  unsigned n_threads = 1;
  // Compresses all full blocks of `pending` (BGZF blocks are independent: one worker per slice of blocks) and writes them
  // in order; with `all` also the trailing partial block.
  bool flush(bool all) {
    const size_t n_blocks = all ? (pending.size() + BLOCK - 1) / BLOCK : pending.size() / BLOCK;
    if (n_blocks == 0) return true;
    const size_t stride = 65536 + 1024;
    std::vector<uint8_t> out(n_blocks * stride);
    std::vector<size_t> out_len(n_blocks, 0);
    auto work = [&](size_t b0, size_t b1) {
      for (size_t b = b0; b < b1; ++b) {
        const size_t o = b * BLOCK, n = std::min(BLOCK, pending.size() - o);
        out_len[b] = bgzf_compress(pending.data() + o, n, out.data() + b * stride, level);
      }
    };
    const size_t nt = std::min<size_t>(n_threads, n_blocks);
    if (nt <= 1) {
      work(0, n_blocks);
    } else {
      std::vector<std::thread> th;
      for (size_t t = 0; t < nt; ++t) th.emplace_back(work, n_blocks * t / nt, n_blocks * (t + 1) / nt);
      for (std::thread& t : th) t.join();
    }
    for (size_t b = 0; b < n_blocks; ++b) {
      if (!out_len[b] || fwrite(out.data() + b * stride, 1, out_len[b], f) != out_len[b]) return false;
    }
    const size_t consumed = std::min(pending.size(), n_blocks * BLOCK);
    pending.erase(pending.begin(), pending.begin() + consumed);
    return true;
  }
  bool write(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    pending.insert(pending.end(), b, b + n);
    return true;
  }
  bool finish() {
    if (!flush(true)) return false;
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    return fwrite(eof, 1, 28, f) == 28;
  }
};

void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put16(std::vector<uint8_t>& v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }

int reg2bin(int64_t beg, int64_t end) {  // SAM spec 5.3
  --end;
  if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

uint8_t nt16(uint8_t b) {
  static const char* tab = "=ACMGRSVTWYHKDBN";
  const char* p = strchr(tab, b);
  return p && b ? (uint8_t)(p - tab) : 15;
}
void tag_z(std::vector<uint8_t>& v, const char* tag, const char* s, size_t n) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('Z'); v.insert(v.end(), s, s + n); v.push_back(0); }
void tag_i(std::vector<uint8_t>& v, const char* tag, int32_t x) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('i'); put32(v, (uint32_t)x); }
void tag_f(std::vector<uint8_t>& v, const char* tag, float x) { uint32_t u; memcpy(&u, &x, 4); v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('f'); put32(v, u); }
// IUPAC-aware complement (bio::alphabets::dna::complement, used by dna::revcomp in record.rs:161)
uint8_t comp_iupac(uint8_t b) {
  static const char* from = "ACGTRYKMBDHVNSWacgtrykmbdhvnsw";
  static const char* to = "TGCAYRMKVHDBNSWtgcayrmkvhdbnsw";
  const char* p = b ? strchr(from, b) : nullptr;
  return p ? (uint8_t)to[p - from] : b;
}

bool read_exact(gzFile f, void* dst, size_t n) {
  uint8_t* d = (uint8_t*)dst;
  while (n) {
    int got = gzread(f, d, (unsigned)std::min<size_t>(n, 1u << 30));
    if (got <= 0) return false;
    d += got; n -= (size_t)got;
  }
  return true;
}

// Length of one BAM auxiliary field starting at p (tag, type, value), 0 if malformed / truncated.
size_t aux_field_len(const uint8_t* p, size_t n) {
  if (n < 3) return 0;
  size_t need;
  switch (p[2]) {
    case 'A': case 'c': case 'C': need = 4; break;
    case 's': case 'S': need = 5; break;
    case 'i': case 'I': case 'f': need = 7; break;
    case 'Z': case 'H': {
      const void* z = memchr(p + 3, 0, n - 3);
      if (!z) return 0;
      need = (size_t)((const uint8_t*)z - p) + 1;
      break;
    }
    case 'B': {
      if (n < 8) return 0;
      uint32_t cnt;
      memcpy(&cnt, p + 4, 4);
      size_t es;
      switch (p[3]) { case 'c': case 'C': es = 1; break; case 's': case 'S': es = 2; break; case 'i': case 'I': case 'f': es = 4; break; default: return 0; }
      need = 8 + (size_t)cnt * es;
      break;
    }
    default: return 0;
  }
  return need <= n ? need : 0;
}

// Input tags that create_bam_record drops (mapping.rs:834-846): BWA / mapAD specific ones, and RG if one is given.
bool aux_dropped(const uint8_t* tag, bool have_read_group) {
  static const char* filter[] = {"AS", "MD", "NM", "X0", "X1", "XA", "XD", "XE", "XF", "XG", "XM", "XN", "XO", "XS", "XT"};
  for (const char* t : filter) if (tag[0] == (uint8_t)t[0] && tag[1] == (uint8_t)t[1]) return true;
  return have_read_group && tag[0] == 'R' && tag[1] == 'G';
}

void tag_a(std::vector<uint8_t>& v, const char* tag, char c) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('A'); v.push_back((uint8_t)c); }

}  // namespace

extern "C" {

// ---- FASTQ ----------------------------------------------------------------------------------------------------
// Opens FASTQ, FASTQ.GZ or BAM input; the format is sniffed from the (decompressed) first bytes.
int mapad_input_open(const char* path, void** out) {
  if (!path || !out) return MAPAD_EINVAL;
  *out = nullptr;
  InputReader* r = new (std::nothrow) InputReader();
  if (!r) return MAPAD_ENOMEM;
  r->f = gzopen(path, "rb");
  if (!r->f) { delete r; return MAPAD_EIO; }
  gzbuffer(r->f, 1 << 20);
  char magic[4] = {0, 0, 0, 0};
  const int got = gzread(r->f, magic, 4);
  auto fail = [&](int rc) { gzclose(r->f); delete r; return rc; };
  if (got == 4 && memcmp(magic, "BAM\1", 4) == 0) {
    r->is_bam = true;
    uint32_t l_text, n_ref;
    if (!read_exact(r->f, &l_text, 4)) return fail(MAPAD_EIO);
    r->header_text.resize(l_text);
    if (l_text && !read_exact(r->f, &r->header_text[0], l_text)) return fail(MAPAD_EIO);
    while (!r->header_text.empty() && r->header_text.back() == 0) r->header_text.pop_back();
    if (!read_exact(r->f, &n_ref, 4)) return fail(MAPAD_EIO);
    for (uint32_t i = 0; i < n_ref; ++i) {  // the input's reference dictionary is not used (@SQ comes from the index)
      uint32_t l_name;
      if (!read_exact(r->f, &l_name, 4)) return fail(MAPAD_EIO);
      std::string skip((size_t)l_name + 4, '\0');
      if (!read_exact(r->f, &skip[0], skip.size())) return fail(MAPAD_EIO);
    }
  } else if (got == 4 && memcmp(magic, "CRAM", 4) == 0) {
    return fail(MAPAD_EINVAL);  // CRAM input is not supported
  } else if (got > 0) {
    r->prefix.assign(magic, magic + got);
  }
  *out = r;
  return MAPAD_OK;
}
int mapad_input_is_bam(void* reader) { return reader && ((InputReader*)reader)->is_bam ? 1 : 0; }
// SAM header text of a BAM input (NULL for FASTQ); valid until the reader is closed.
const char* mapad_input_header_text(void* reader) {
  InputReader* r = (InputReader*)reader;
  return r && r->is_bam ? r->header_text.c_str() : nullptr;
}

// BAM records -> Record (record.rs:138-182).
static int bam_next_chunk(InputReader* r, uint64_t max_reads, ReadChunk* c) {
  std::vector<uint8_t> blk;
  while ((uint64_t)c->flags.size() < max_reads) {
    uint32_t block_size;
    const int got = gzread(r->f, &block_size, 4);
    if (got == 0) break;  // end of file
    if (got != 4 || block_size < 32) return MAPAD_EIO;
    blk.resize(block_size);
    if (!read_exact(r->f, blk.data(), block_size)) return MAPAD_EIO;
    const uint8_t l_read_name = blk[8];
    uint16_t n_cigar, flag;
    uint32_t l_seq;
    memcpy(&n_cigar, &blk[12], 2); memcpy(&flag, &blk[14], 2); memcpy(&l_seq, &blk[16], 4);
    size_t o = 32;
    const size_t need = o + l_read_name + 4ull * n_cigar + (l_seq + 1ull) / 2 + l_seq;
    if (need > block_size) return MAPAD_EIO;
    const char* name = (const char*)&blk[o];
    size_t name_len = l_read_name ? strnlen(name, l_read_name) : 0;
    if (name_len == 1 && name[0] == '*') name_len = 0;  // missing name
    o += l_read_name + 4ull * n_cigar;
    const uint8_t* sq = &blk[o];
    o += (l_seq + 1ull) / 2;
    const uint8_t* ql = &blk[o];
    o += l_seq;
    // missing qualities (0xff) leave sequence and quality lengths different: skipped (input_chunk_reader.rs:206-214);
    // reads longer than i16::MAX are an error there (record.rs:143-149), skipped here
    if ((l_seq > 0 && ql[0] == 0xff) || l_seq > 32767) { c->skipped += 1; continue; }
    static const char* dec = "=ACMGRSVTWYHKDBN";
    const bool rev = (flag & 0x10) != 0;  // stored reverse-complemented: turn back (record.rs:159-162)
    const size_t base = c->seq.size();
    c->seq.resize(base + l_seq);
    c->qual.resize(base + l_seq);
    for (uint32_t i = 0; i < l_seq; ++i) {
      const uint8_t b = (uint8_t)dec[(sq[i >> 1] >> ((~i & 1) << 2)) & 15];
      if (rev) { c->seq[base + l_seq - 1 - i] = comp_iupac(b); c->qual[base + l_seq - 1 - i] = ql[i]; }
      else { c->seq[base + i] = b; c->qual[base + i] = ql[i]; }
    }
    c->offsets.push_back(c->seq.size());
    c->names.insert(c->names.end(), name, name + name_len);
    c->name_offsets.push_back(c->names.size());
    c->flags.push_back(flag);
    // auxiliary fields are kept verbatim (typed copies in the reference, record.rs:164-172); validate their framing
    size_t a = o;
    while (a < block_size) {
      const size_t fl = aux_field_len(&blk[a], block_size - a);
      if (!fl) return MAPAD_EIO;
      a += fl;
    }
    c->aux.insert(c->aux.end(), blk.begin() + o, blk.end());
    c->aux_offsets.push_back(c->aux.size());
  }
  return MAPAD_OK;
}

// Reads up to `max_reads` records (the reference's --batch_size chunking, input_chunk_reader.rs:176-244); records whose
// sequence and quality lengths differ or that are longer than i16::MAX are skipped like there (:200-214, record.rs:188).
int mapad_input_next_chunk(void* reader, uint64_t max_reads, void** chunk_out) {
  if (!reader || !chunk_out) return MAPAD_EINVAL;
  InputReader* r = (InputReader*)reader;
  ReadChunk* c = new (std::nothrow) ReadChunk();
  if (!c) return MAPAD_ENOMEM;
  if (r->is_bam) {
    int rc;
    try { rc = bam_next_chunk(r, max_reads, c); } catch (const std::bad_alloc&) { rc = MAPAD_ENOMEM; }
    if (rc != MAPAD_OK) { delete c; return rc; }
    *chunk_out = c;
    return MAPAD_OK;
  }
  std::string name, seq, qual;
  while ((uint64_t)c->flags.size() < max_reads) {
    if (!r->getline()) break;
    if (r->line.empty()) continue;
    if (r->line[0] != '@') { delete c; return MAPAD_EIO; }
    name = r->line.substr(1);
    size_t sp = name.find_first_of(" \t");
    if (sp != std::string::npos) name.resize(sp);  // noodles: name = up to the first whitespace, rest is the description
    if (!r->getline()) { delete c; return MAPAD_EIO; }
    seq = r->line;
    if (!r->getline() || r->line.empty() || r->line[0] != '+') { delete c; return MAPAD_EIO; }
    if (!r->getline()) { delete c; return MAPAD_EIO; }
    qual = r->line;
    if (seq.size() != qual.size() || seq.size() > 32767) { c->skipped += 1; continue; }
    for (char ch : seq) c->seq.push_back((uint8_t)((ch >= 'a' && ch <= 'z') ? ch - 32 : ch));
    for (char ch : qual) c->qual.push_back((uint8_t)(ch - 33));
    c->offsets.push_back(c->seq.size());
    c->names.insert(c->names.end(), name.begin(), name.end());
    c->name_offsets.push_back(c->names.size());
    c->flags.push_back(0);
    c->aux_offsets.push_back(0);
  }
  *chunk_out = c;
  return MAPAD_OK;
}
void mapad_input_close(void* reader) {
  InputReader* r = (InputReader*)reader;
  if (r) { if (r->f) gzclose(r->f); delete r; }
}
// Raw BAM auxiliary fields of the chunk's reads: read i owns aux[aux_offsets[i] .. aux_offsets[i+1]).
int mapad_chunk_aux(void* chunk, const uint8_t** aux, const uint64_t** aux_offsets) {
  ReadChunk* c = (ReadChunk*)chunk;
  if (!c || !aux || !aux_offsets) return MAPAD_EINVAL;
  *aux = c->aux.data();
  *aux_offsets = c->aux_offsets.data();
  return MAPAD_OK;
}
// Views into a chunk (valid until mapad_chunk_free).
uint64_t mapad_chunk_view(void* chunk, mapad_reads* reads, const char** names, const uint64_t** name_offsets, const uint16_t** flags,
                          uint64_t* skipped) {
  ReadChunk* c = (ReadChunk*)chunk;
  if (!c) return 0;
  if (reads) {
    memset(reads, 0, sizeof *reads);
    reads->n_reads = c->flags.size(); reads->seq = c->seq.data(); reads->qual = c->qual.data(); reads->offsets = c->offsets.data();
  }
  if (names) *names = c->names.data();
  if (name_offsets) *name_offsets = c->name_offsets.data();
  if (flags) *flags = c->flags.data();
  if (skipped) *skipped = c->skipped;
  return c->flags.size();
}
void mapad_chunk_free(void* chunk) { delete (ReadChunk*)chunk; }

// ---- BAM ------------------------------------------------------------------------------------------------------
// create_bam_header (mapping.rs:300-398): @HD VN:1.6 SO:unsorted, one @SQ per contig of the index, @RG (the one given on
// the command line, else those of the input header), the input's @PG chain plus a new @PG (ID made unique as
// "mapAD.<n>", PP = the chain's leaf — noodles `Programs::add`), the input's @CO lines.  `src_header_text` is the SAM
// header of a BAM input or NULL.
int mapad_bam_open_with_header(const char* path, const mapad_index* index, const char* command_line, const char* read_group_id,
                               int force_overwrite, const char* src_header_text, void** out) {
  if (!path || !index || !out) return MAPAD_EINVAL;
  const mapad::HostIndex* ix = reinterpret_cast<const mapad::HostIndex*>(index);
  if (!force_overwrite) {  // OpenOptions::create_new (mapping.rs:92-100)
    FILE* t = fopen(path, "rb");
    if (t) { fclose(t); return MAPAD_EIO; }
  }
  BamWriter* w = new (std::nothrow) BamWriter();
  if (!w) return MAPAD_ENOMEM;
  w->f = fopen(path, "wb");
  if (!w->f) { delete w; return MAPAD_EIO; }
  w->contig_names = ix->contig_names;
  if (read_group_id) w->read_group = read_group_id;
  {  // BGZF blocks are compressed by a few host threads (MAPAD_BAM_THREADS, default min(hardware threads, 16))
    const char* e = getenv("MAPAD_BAM_THREADS");
    unsigned hw = std::thread::hardware_concurrency();
    w->n_threads = e ? (unsigned)std::max(1, atoi(e)) : std::max(1u, std::min(hw ? hw : 1u, 16u));
    if (const char* l = getenv("MAPAD_BAM_LEVEL")) w->level = std::min(9, std::max(0, atoi(l)));
  }
  // pick @RG / @PG / @CO lines out of the input header
  std::vector<std::string> src_rg, src_pg, src_co;
  if (src_header_text) {
    const std::string src(src_header_text);
    size_t b = 0;
    while (b < src.size()) {
      size_t e = src.find('\n', b);
      if (e == std::string::npos) e = src.size();
      std::string line = src.substr(b, e - b);
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.compare(0, 3, "@RG") == 0) src_rg.push_back(line);
      else if (line.compare(0, 3, "@PG") == 0) src_pg.push_back(line);
      else if (line.compare(0, 3, "@CO") == 0) src_co.push_back(line);
      b = e + 1;
    }
  }
  auto field = [](const std::string& line, const char* key) -> std::string {  // value of "\tKEY:" in a header line
    const std::string k = std::string("\t") + key + ":";
    size_t p = line.find(k);
    if (p == std::string::npos) return "";
    p += k.size();
    size_t e = line.find('\t', p);
    return line.substr(p, e == std::string::npos ? std::string::npos : e - p);
  };
  std::string text = "@HD\tVN:1.6\tSO:unsorted\n";
  for (size_t i = 0; i < ix->contig_names.size(); ++i)
    text += "@SQ\tSN:" + ix->contig_names[i] + "\tLN:" + std::to_string(ix->contig_end[i] - ix->contig_start[i] + 1) + "\n";
  if (!w->read_group.empty()) text += "@RG\tID:" + w->read_group + "\n";
  else for (const std::string& l : src_rg) text += l + "\n";
  std::vector<std::string> pg_ids, pg_pps;
  for (const std::string& l : src_pg) { text += l + "\n"; pg_ids.push_back(field(l, "ID")); pg_pps.push_back(field(l, "PP")); }
  std::string program_id = "mapAD";
  {
    size_t taken = 0;
    for (const std::string& id : pg_ids) if (id == program_id || id.compare(0, program_id.size() + 1, program_id + ".") == 0) taken += 1;
    if (taken > 0) program_id += "." + std::to_string(taken);
  }
  const std::string pg_body = std::string("\tPN:mapAD\tVN:0.45.0-b200\tDS:An aDNA aware short-read mapper\tCL:") + (command_line ? command_line : "");
  if (pg_ids.empty()) {
    text += "@PG\tID:" + program_id + pg_body + "\n";
  } else {  // one new entry per chain leaf
    std::vector<std::string> used = pg_ids;
    for (size_t i = 0; i < pg_ids.size(); ++i) {
      if (std::find(pg_pps.begin(), pg_pps.end(), pg_ids[i]) != pg_pps.end()) continue;  // not a leaf
      std::string id = program_id;
      if (std::find(used.begin(), used.end(), id) != used.end()) id += "-" + pg_ids[i];
      used.push_back(id);
      text += "@PG\tID:" + id + pg_body + "\tPP:" + pg_ids[i] + "\n";
    }
  }
  for (const std::string& l : src_co) text += l + "\n";
  std::vector<uint8_t> h;
  h.insert(h.end(), {'B', 'A', 'M', 1});
  put32(h, (uint32_t)text.size());
  h.insert(h.end(), text.begin(), text.end());
  put32(h, (uint32_t)ix->contig_names.size());
  for (size_t i = 0; i < ix->contig_names.size(); ++i) {
    put32(h, (uint32_t)ix->contig_names[i].size() + 1);
    h.insert(h.end(), ix->contig_names[i].begin(), ix->contig_names[i].end());
    h.push_back(0);
    put32(h, (uint32_t)(ix->contig_end[i] - ix->contig_start[i] + 1));
  }
  if (!w->write(h.data(), h.size()) || !w->flush(false)) { fclose(w->f); delete w; return MAPAD_EIO; }
  *out = w;
  return MAPAD_OK;
}
int mapad_bam_open(const char* path, const mapad_index* index, const char* command_line, const char* read_group_id, int force_overwrite,
                   void** out) {
  return mapad_bam_open_with_header(path, index, command_line, read_group_id, force_overwrite, nullptr, out);
}

// create_bam_record (mapping.rs:722-927) for every read of a chunk, in input order.
// `aux` / `aux_offsets` (may be NULL): the input records' auxiliary fields, copied except for the filtered tags.
int mapad_bam_write_chunk_aux(void* writer, const mapad_index* index, const mapad_reads* reads, const char* names,
                              const uint64_t* name_offsets, const uint16_t* in_flags, const uint8_t* aux, const uint64_t* aux_offsets,
                              const mapad_results* res) {
  if (!writer || !index || !reads || !res || reads->n_reads != res->n_reads) return MAPAD_EINVAL;
  BamWriter* w = (BamWriter*)writer;
  std::vector<uint8_t> rec, tags;
  std::vector<char> xa(1 << 16);
  for (uint64_t r = 0; r < reads->n_reads; ++r) {
    const mapad_record& m = res->records[r];
    const uint64_t o = reads->offsets[r], L = reads->offsets[r + 1] - o;
    const bool has_name = names && name_offsets[r + 1] > name_offsets[r];  // a missing name is written as "*"
    const char* nm = has_name ? names + name_offsets[r] : "*";
    const size_t nml = has_name ? (size_t)(name_offsets[r + 1] - name_offsets[r]) : 1;
    if (nml > 254) return MAPAD_EINVAL;  // l_read_name is one byte incl. the NUL (noodles rejects such names as well)
    uint16_t flag = in_flags ? in_flags[r] : 0;
    flag &= (uint16_t)~(0x8 | 0x20 | 0x2 | 0x100 | 0x800);                 // :750-755
    if (m.mapped) flag &= (uint16_t)~0x4; else { flag |= 0x4; flag &= (uint16_t)~(0x10 | 0x2); }  // :757-769
    if (m.mapped && m.strand) flag |= 0x10; else flag &= (uint16_t)~0x10;  // :771-776
    const int32_t ref_id = m.mapped ? m.tid : -1;
    const int32_t pos = m.mapped ? (int32_t)m.pos : -1;
    int64_t ref_span = 0;
    for (uint32_t i = 0; i < m.cigar_len; ++i) { uint32_t v = res->cigar[m.cigar_off + i]; if ((v & 15) != 1) ref_span += v >> 4; }
    const int bin = m.mapped ? reg2bin(pos, pos + (ref_span > 0 ? ref_span : 1)) : 4680;
    rec.clear();
    put32(rec, 0);  // block_size, patched below
    put32(rec, (uint32_t)ref_id);
    put32(rec, (uint32_t)pos);
    rec.push_back((uint8_t)(nml + 1));
    rec.push_back((uint8_t)m.mapq);
    put16(rec, (uint16_t)bin);
    put16(rec, (uint16_t)(m.mapped ? m.cigar_len : 0));
    put16(rec, flag);
    put32(rec, (uint32_t)L);
    put32(rec, (uint32_t)-1); put32(rec, (uint32_t)-1); put32(rec, 0);  // mate reference / position, template length
    rec.insert(rec.end(), nm, nm + nml); rec.push_back(0);
    if (m.mapped) for (uint32_t i = 0; i < m.cigar_len; ++i) put32(rec, res->cigar[m.cigar_off + i]);
    // sequence and qualities: reversed for reverse-strand hits (:795-819)
    const bool rev = m.mapped && m.strand;
    for (uint64_t i = 0; i < L; i += 2) {
      auto base = [&](uint64_t k) -> uint8_t { return rev ? comp_iupac(reads->seq[o + L - 1 - k]) : reads->seq[o + k]; };
      uint8_t hi = nt16(base(i)), lo = i + 1 < L ? nt16(base(i + 1)) : 0;
      rec.push_back((uint8_t)(hi << 4 | lo));
    }
    for (uint64_t i = 0; i < L; ++i) rec.push_back(rev ? reads->qual[o + L - 1 - i] : reads->qual[o + i]);
    // tags (:829-918): the input's fields minus the filter list, then ours
    tags.clear();
    if (aux && aux_offsets) {
      const uint8_t* a = aux + aux_offsets[r];
      size_t left = (size_t)(aux_offsets[r + 1] - aux_offsets[r]);
      while (left) {
        const size_t fl = aux_field_len(a, left);
        if (!fl) return MAPAD_EINVAL;
        if (!aux_dropped(a, !w->read_group.empty())) tags.insert(tags.end(), a, a + fl);
        a += fl; left -= fl;
      }
    }
    if (!w->read_group.empty()) tag_z(tags, "RG", w->read_group.data(), w->read_group.size());
    if (m.mapped) {
      tag_f(tags, "AS", m.alignment_score);
      tag_i(tags, "NM", m.nm);
      tag_z(tags, "MD", res->text + m.md_off, m.md_len);
      if (m.n_alts) {
        int64_t k = mapad_format_xa(index, res, r, xa.data(), xa.size());
        if (k > 0) tag_z(tags, "XA", xa.data(), (size_t)k);
      }
      tag_i(tags, "X0", m.x0);
      tag_i(tags, "X1", m.x1);
      if (m.x1 > 0) tag_f(tags, "XS", m.xs);
      tag_a(tags, "XT", (char)m.xt);
    }
    rec.insert(rec.end(), tags.begin(), tags.end());
    const uint32_t bs = (uint32_t)rec.size() - 4;
    memcpy(rec.data(), &bs, 4);
    if (!w->write(rec.data(), rec.size())) return MAPAD_EIO;
  }
  return w->flush(false) ? MAPAD_OK : MAPAD_EIO;
}

int mapad_bam_write_chunk(void* writer, const mapad_index* index, const mapad_reads* reads, const char* names, const uint64_t* name_offsets,
                          const uint16_t* in_flags, const mapad_results* res) {
  return mapad_bam_write_chunk_aux(writer, index, reads, names, name_offsets, in_flags, nullptr, nullptr, res);
}

int mapad_bam_close(void* writer) {
  BamWriter* w = (BamWriter*)writer;
  if (!w) return MAPAD_EINVAL;
  bool ok = w->finish();
  ok = (fclose(w->f) == 0) && ok;
  delete w;
  return ok ? MAPAD_OK : MAPAD_EIO;
}

}  // extern "C"
