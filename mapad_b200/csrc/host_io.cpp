// host_io.cpp — callers and data formats on either side of the hot path:
//   * FASTQ / FASTQ.GZ reader with the reference's Record normalisation
//       (/root/reference/src/map/record.rs:184-215: upper-case sequence, Phred+33 removed, flags 0)
//   * BAM writer: create_bam_header (src/map/mapping.rs:300-398) and create_bam_record (:722-927) on top of the
//     per-read fields the device produced, BGZF-compressed with zlib.
// The reference uses the noodles crate for both; only the byte formats (SAM/BAM spec) are shared.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
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
  uint64_t skipped = 0;
};

struct FastqReader {
  gzFile f = nullptr;  // gzopen reads plain files transparently
  std::string line;
  bool getline() {
    line.clear();
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
struct BamWriter {
  FILE* f = nullptr;
  std::vector<uint8_t> pending;
  std::vector<std::string> contig_names;
  std::string read_group;  // ID or empty
  bool flush_block(const uint8_t* data, size_t n) {
    uint8_t out[65536 + 1024];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    zs.next_in = const_cast<uint8_t*>(data);
    zs.avail_in = (uInt)n;
    zs.next_out = out + 18;
    zs.avail_out = sizeof out - 18 - 8;
    int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) { deflateEnd(&zs); return false; }
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const size_t bsize = clen + 18 + 8;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    out[16] = (uint8_t)((bsize - 1) & 0xff); out[17] = (uint8_t)((bsize - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n);
    const uint32_t isize = (uint32_t)n;
    memcpy(out + 18 + clen, &crc, 4);
    memcpy(out + 18 + clen + 4, &isize, 4);
    return fwrite(out, 1, bsize, f) == bsize;
  }
  bool write(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    pending.insert(pending.end(), b, b + n);
    while (pending.size() >= 0xff00) {
      if (!flush_block(pending.data(), 0xff00)) return false;
      pending.erase(pending.begin(), pending.begin() + 0xff00);
    }
    return true;
  }
  bool finish() {
    if (!pending.empty() && !flush_block(pending.data(), pending.size())) return false;
    pending.clear();
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

uint8_t comp_base(uint8_t b) {
  switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return b; }
}
uint8_t nt16(uint8_t b) {
  static const char* tab = "=ACMGRSVTWYHKDBN";
  const char* p = strchr(tab, b);
  return p && b ? (uint8_t)(p - tab) : 15;
}
void tag_z(std::vector<uint8_t>& v, const char* tag, const char* s, size_t n) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('Z'); v.insert(v.end(), s, s + n); v.push_back(0); }
void tag_i(std::vector<uint8_t>& v, const char* tag, int32_t x) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('i'); put32(v, (uint32_t)x); }
void tag_f(std::vector<uint8_t>& v, const char* tag, float x) { uint32_t u; memcpy(&u, &x, 4); v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('f'); put32(v, u); }
void tag_a(std::vector<uint8_t>& v, const char* tag, char c) { v.push_back(tag[0]); v.push_back(tag[1]); v.push_back('A'); v.push_back((uint8_t)c); }

}  // namespace

extern "C" {

// ---- FASTQ ----------------------------------------------------------------------------------------------------
int mapad_fastq_open(const char* path, void** out) {
  if (!path || !out) return MAPAD_EINVAL;
  FastqReader* r = new (std::nothrow) FastqReader();
  if (!r) return MAPAD_ENOMEM;
  r->f = gzopen(path, "rb");
  if (!r->f) { delete r; return MAPAD_EIO; }
  gzbuffer(r->f, 1 << 20);
  *out = r;
  return MAPAD_OK;
}
// Reads up to `max_reads` records (the reference's --batch_size chunking, input_chunk_reader.rs:176-244); records whose
// sequence and quality lengths differ or that are longer than i16::MAX are skipped like there (:200-214, record.rs:188).
int mapad_fastq_next_chunk(void* reader, uint64_t max_reads, void** chunk_out) {
  if (!reader || !chunk_out) return MAPAD_EINVAL;
  FastqReader* r = (FastqReader*)reader;
  ReadChunk* c = new (std::nothrow) ReadChunk();
  if (!c) return MAPAD_ENOMEM;
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
  }
  *chunk_out = c;
  return MAPAD_OK;
}
void mapad_fastq_close(void* reader) {
  FastqReader* r = (FastqReader*)reader;
  if (r) { if (r->f) gzclose(r->f); delete r; }
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
// create_bam_header (mapping.rs:300-398) for FASTQ input: @HD VN:1.6 SO:unsorted, one @SQ per contig, optional @RG,
// @PG ID:mapAD with the command line.
int mapad_bam_open(const char* path, const mapad_index* index, const char* command_line, const char* read_group_id, int force_overwrite,
                   void** out) {
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
  std::string text = "@HD\tVN:1.6\tSO:unsorted\n";
  for (size_t i = 0; i < ix->contig_names.size(); ++i)
    text += "@SQ\tSN:" + ix->contig_names[i] + "\tLN:" + std::to_string(ix->contig_end[i] - ix->contig_start[i] + 1) + "\n";
  if (!w->read_group.empty()) text += "@RG\tID:" + w->read_group + "\n";
  text += std::string("@PG\tID:mapAD\tPN:mapAD\tVN:0.45.0-b200\tDS:An aDNA aware short-read mapper\tCL:") + (command_line ? command_line : "") + "\n";
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
  if (!w->write(h.data(), h.size())) { fclose(w->f); delete w; return MAPAD_EIO; }
  *out = w;
  return MAPAD_OK;
}

// create_bam_record (mapping.rs:722-927) for every read of a chunk, in input order.
int mapad_bam_write_chunk(void* writer, const mapad_index* index, const mapad_reads* reads, const char* names, const uint64_t* name_offsets,
                          const uint16_t* in_flags, const mapad_results* res) {
  if (!writer || !index || !reads || !res || reads->n_reads != res->n_reads) return MAPAD_EINVAL;
  BamWriter* w = (BamWriter*)writer;
  std::vector<uint8_t> rec, tags;
  std::vector<char> xa(1 << 16);
  for (uint64_t r = 0; r < reads->n_reads; ++r) {
    const mapad_record& m = res->records[r];
    const uint64_t o = reads->offsets[r], L = reads->offsets[r + 1] - o;
    const char* nm = names ? names + name_offsets[r] : "*";
    const size_t nml = names ? (size_t)(name_offsets[r + 1] - name_offsets[r]) : 1;
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
      auto base = [&](uint64_t k) -> uint8_t { return rev ? comp_base(reads->seq[o + L - 1 - k]) : reads->seq[o + k]; };
      uint8_t hi = nt16(base(i)), lo = i + 1 < L ? nt16(base(i + 1)) : 0;
      rec.push_back((uint8_t)(hi << 4 | lo));
    }
    for (uint64_t i = 0; i < L; ++i) rec.push_back(rev ? reads->qual[o + L - 1 - i] : reads->qual[o + i]);
    // tags (:850-918); FASTQ input carries none to copy
    tags.clear();
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
  return MAPAD_OK;
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
