// host_index_files.cpp — the reference's on-disk index: seven files
//   <ref>.tbw BWT, .tle Less, .toc Occ, .trt RankTransform, .tsa sampled suffix array, .tpi contig map, .tos original symbols
// each a Snappy *frame* stream (snap::write::FrameEncoder, /root/reference/src/index/indexing.rs:111-207; read with
// FrameDecoder, src/index/versioned_index.rs:52-54) around a bincode-1.3 payload `Item{version: u8 = 5, data}`
// (versioned_index.rs:12-19).  The writer emits uncompressed frame chunks (valid Snappy framing); the reader also
// decodes compressed chunks.  .tsa/.tpi/.tos layouts are fully defined by the reference (src/index/mod.rs:32-42,80-86,
// 198-199); .tbw/.tle/.toc/.trt serialise rust-bio types (Vec<u8>, Vec<usize>, Occ{occ: Vec<Vec<usize>>, k: u32},
// RankTransform{ranks: VecMap<u8>}) whose byte layout could not be checked against files written by mapAD itself.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/mapad_gpu.h"
#include "host_index.hpp"

namespace {

const uint8_t INDEX_VERSION = 5;

// ---- CRC-32C (Castagnoli), masked as in the Snappy framing format -------------------------------------------
uint32_t crc32c_table[256];
bool crc_init_done = false;
void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    crc32c_table[i] = c;
  }
  crc_init_done = true;
}
uint32_t crc32c(const uint8_t* p, size_t n) {
  if (!crc_init_done) crc_init();
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) c = crc32c_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}
uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

// ---- frame writer ---------------------------------------------------------------------------------------------
struct FrameWriter {
  FILE* f = nullptr;
  std::vector<uint8_t> buf;
  bool ok = true;
  bool open(const std::string& path) {
    f = fopen(path.c_str(), "wb");
    if (!f) return false;
    static const uint8_t ident[10] = {0xff, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
    ok = fwrite(ident, 1, 10, f) == 10;
    return ok;
  }
  void flush_chunk() {
    if (buf.empty()) return;
    const uint32_t len = (uint32_t)buf.size() + 4;
    uint8_t hdr[8] = {0x01, (uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), 0, 0, 0, 0};
    const uint32_t c = mask_crc(crc32c(buf.data(), buf.size()));
    memcpy(hdr + 4, &c, 4);
    ok = ok && fwrite(hdr, 1, 8, f) == 8 && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    buf.clear();
  }
  void write(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    while (n) {
      size_t take = std::min<size_t>(n, 65536 - buf.size());
      buf.insert(buf.end(), b, b + take);
      b += take; n -= take;
      if (buf.size() == 65536) flush_chunk();
    }
  }
  void u8(uint8_t v) { write(&v, 1); }
  void u32(uint32_t v) { write(&v, 4); }
  void u64(uint64_t v) { write(&v, 8); }
  bool close() { flush_chunk(); bool r = ok && fclose(f) == 0; f = nullptr; return r; }
};

// ---- raw Snappy block decoder -----------------------------------------------------------------------------------
bool snappy_uncompress(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
  size_t ip = 0;
  uint64_t ulen = 0;
  int shift = 0;
  while (true) {
    if (ip >= n || shift > 35) return false;
    uint8_t b = in[ip++];
    ulen |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
  }
  out.clear();
  out.reserve(ulen);
  while (ip < n) {
    const uint8_t tag = in[ip++];
    const int type = tag & 3;
    if (type == 0) {  // literal
      size_t len = (tag >> 2) + 1;
      if (len > 60) {
        const int nb = (int)len - 60;
        if (ip + nb > n) return false;
        len = 0;
        for (int k = 0; k < nb; ++k) len |= (size_t)in[ip + k] << (8 * k);
        len += 1;
        ip += nb;
      }
      if (ip + len > n) return false;
      out.insert(out.end(), in + ip, in + ip + len);
      ip += len;
    } else {
      size_t len, off;
      if (type == 1) {
        if (ip + 1 > n) return false;
        len = ((tag >> 2) & 7) + 4;
        off = ((size_t)(tag >> 5) << 8) | in[ip];
        ip += 1;
      } else if (type == 2) {
        if (ip + 2 > n) return false;
        len = (tag >> 2) + 1;
        off = in[ip] | ((size_t)in[ip + 1] << 8);
        ip += 2;
      } else {
        if (ip + 4 > n) return false;
        len = (tag >> 2) + 1;
        off = in[ip] | ((size_t)in[ip + 1] << 8) | ((size_t)in[ip + 2] << 16) | ((size_t)in[ip + 3] << 24);
        ip += 4;
      }
      if (off == 0 || off > out.size()) return false;
      size_t start = out.size() - off;
      for (size_t k = 0; k < len; ++k) out.push_back(out[start + k]);  // may overlap
    }
  }
  return out.size() == ulen;
}

// ---- frame reader -----------------------------------------------------------------------------------------------
struct FrameReader {
  FILE* f = nullptr;
  std::vector<uint8_t> chunk, raw;
  size_t pos = 0;
  bool ok = true;
  bool open(const std::string& path) { f = fopen(path.c_str(), "rb"); return f != nullptr; }
  bool next_chunk() {
    while (true) {
      uint8_t hdr[4];
      size_t got = fread(hdr, 1, 4, f);
      if (got == 0) return false;
      if (got != 4) { ok = false; return false; }
      const uint32_t len = hdr[1] | (hdr[2] << 8) | (hdr[3] << 16);
      raw.resize(len);
      if (len && fread(raw.data(), 1, len, f) != len) { ok = false; return false; }
      const uint8_t type = hdr[0];
      if (type == 0xff) { if (len != 6 || memcmp(raw.data(), "sNaPpY", 6) != 0) { ok = false; return false; } continue; }
      if (type >= 0x80) continue;                       // skippable / padding
      if (type > 0x01 || len < 4) { ok = false; return false; }  // reserved unskippable
      uint32_t want;
      memcpy(&want, raw.data(), 4);
      if (type == 0x01) chunk.assign(raw.begin() + 4, raw.end());
      else if (!snappy_uncompress(raw.data() + 4, len - 4, chunk)) { ok = false; return false; }
      if (mask_crc(crc32c(chunk.data(), chunk.size())) != want) { ok = false; return false; }
      pos = 0;
      if (!chunk.empty()) return true;
    }
  }
  bool read(void* dst, size_t n) {
    uint8_t* d = (uint8_t*)dst;
    while (n) {
      if (pos == chunk.size() && !next_chunk()) { ok = false; return false; }
      size_t take = std::min(n, chunk.size() - pos);
      memcpy(d, chunk.data() + pos, take);
      d += take; pos += take; n -= take;
    }
    return true;
  }
  bool u8(uint8_t& v) { return read(&v, 1); }
  bool u32(uint32_t& v) { return read(&v, 4); }
  bool u64(uint64_t& v) { return read(&v, 8); }
  void close() { if (f) fclose(f); f = nullptr; }
};

bool read_version(FrameReader& r, int& rc) {
  uint8_t v;
  if (!r.u8(v)) { rc = MAPAD_EIO; return false; }
  if (v != INDEX_VERSION) { rc = MAPAD_EINDEX; return false; }  // Error::IndexVersionMismatch (versioned_index.rs:36-44)
  return true;
}

}  // namespace

extern "C" {

int mapad_index_save(const mapad_index* index, const char* prefix) {
  if (!index || !prefix) return MAPAD_EINVAL;
  const mapad::HostIndex& ix = *reinterpret_cast<const mapad::HostIndex*>(index);
  const std::string p(prefix);
  const uint64_t n = ix.n;
  {  // .tbw: Item<BWT = Vec<u8>>
    FrameWriter w;
    if (!w.open(p + ".tbw")) return MAPAD_EIO;
    w.u8(INDEX_VERSION); w.u64(n); w.write(ix.bwt.data(), n);
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .tle: Item<Less = Vec<usize>>, max_symbol + 2 = 7 entries for $ACGTX
    FrameWriter w;
    if (!w.open(p + ".tle")) return MAPAD_EIO;
    w.u8(INDEX_VERSION); w.u64(7);
    for (int c = 0; c < 7; ++c) w.u64(ix.less[c]);
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .toc: Item<Occ{occ: Vec<Vec<usize>>, k: u32}>, k = 128 (indexing.rs:188): occ[c][i] = #c in bwt[0..=i*k]
    FrameWriter w;
    if (!w.open(p + ".toc")) return MAPAD_EIO;
    const uint32_t k = 128;
    const uint64_t n_cp = n ? (n - 1) / k + 1 : 0;
    std::vector<std::vector<uint64_t>> occ(6, std::vector<uint64_t>());
    for (auto& v : occ) v.reserve(n_cp);
    uint64_t cur[6] = {0, 0, 0, 0, 0, 0};
    for (uint64_t i = 0; i < n; ++i) {
      cur[ix.bwt[i] < 6 ? ix.bwt[i] : 5] += 1;
      if (i % k == 0) for (int c = 0; c < 6; ++c) occ[c].push_back(cur[c]);
    }
    w.u8(INDEX_VERSION); w.u64(6);
    for (int c = 0; c < 6; ++c) { w.u64(occ[c].size()); w.write(occ[c].data(), occ[c].size() * 8); }
    w.u32(k);
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .trt: Item<RankTransform{ranks: VecMap<u8>}> as a map symbol -> rank over $ACGTX
    FrameWriter w;
    if (!w.open(p + ".trt")) return MAPAD_EIO;
    static const char alpha[6] = {'$', 'A', 'C', 'G', 'T', 'X'};
    w.u8(INDEX_VERSION); w.u64(6);
    for (int r = 0; r < 6; ++r) { w.u64((uint64_t)(uint8_t)alpha[r]); w.u8((uint8_t)r); }
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .tsa: Item<SampledSuffixArrayOwned{sample, sampling_rate, extra_rows, sentinel}> (index/mod.rs:80-86)
    FrameWriter w;
    if (!w.open(p + ".tsa")) return MAPAD_EIO;
    w.u8(INDEX_VERSION); w.u64(ix.sa_sample.size()); w.write(ix.sa_sample.data(), ix.sa_sample.size() * 8);
    w.u64(ix.sa_rate);
    w.u64(ix.extra_rows.size() / 2); w.write(ix.extra_rows.data(), ix.extra_rows.size() * 8);
    w.u8(0);
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .tpi: Item<FastaIdPositions{id_position: Vec<{start, end, identifier}>}> (index/mod.rs:32-42)
    FrameWriter w;
    if (!w.open(p + ".tpi")) return MAPAD_EIO;
    w.u8(INDEX_VERSION); w.u64(ix.contig_start.size());
    for (size_t i = 0; i < ix.contig_start.size(); ++i) {
      w.u64(ix.contig_start[i]); w.u64(ix.contig_end[i]);
      w.u64(ix.contig_names[i].size()); w.write(ix.contig_names[i].data(), ix.contig_names[i].size());
    }
    if (!w.close()) return MAPAD_EIO;
  }
  {  // .tos: Item<OriginalSymbols(BTreeMap<usize, u8>)> (index/mod.rs:198-199)
    FrameWriter w;
    if (!w.open(p + ".tos")) return MAPAD_EIO;
    w.u8(INDEX_VERSION); w.u64(ix.orig_pos.size());
    for (size_t i = 0; i < ix.orig_pos.size(); ++i) { w.u64(ix.orig_pos[i]); w.u8(ix.orig_sym[i]); }
    if (!w.close()) return MAPAD_EIO;
  }
  return MAPAD_OK;
}

int mapad_index_load(const char* prefix, mapad_index** out) {
  if (!prefix || !out) return MAPAD_EINVAL;
  *out = nullptr;
  const std::string p(prefix);
  mapad::HostIndex* ix = new (std::nothrow) mapad::HostIndex();
  if (!ix) return MAPAD_ENOMEM;
  int rc = MAPAD_OK;
  auto fail = [&](int code) { delete ix; return code; };
  try {
    {  // .tbw
      FrameReader r;
      if (!r.open(p + ".tbw")) return fail(MAPAD_EIO);
      uint64_t n;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      if (!r.u64(n)) { r.close(); return fail(MAPAD_EIO); }
      ix->n = n;
      ix->bwt.resize(n);
      bool ok = n == 0 || r.read(ix->bwt.data(), n);
      r.close();
      if (!ok) return fail(MAPAD_EIO);
    }
    {  // .tsa
      FrameReader r;
      if (!r.open(p + ".tsa")) return fail(MAPAD_EIO);
      uint64_t ns, rate, ne;
      uint8_t sentinel;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      bool ok = r.u64(ns);
      if (ok) { ix->sa_sample.resize(ns); ok = ns == 0 || r.read(ix->sa_sample.data(), ns * 8); }
      ok = ok && r.u64(rate) && r.u64(ne);
      if (ok) { ix->sa_rate = rate; ix->extra_rows.resize(2 * ne); ok = ne == 0 || r.read(ix->extra_rows.data(), 16 * ne); }
      ok = ok && r.u8(sentinel);
      r.close();
      if (!ok) return fail(MAPAD_EIO);
      if (sentinel != 0) return fail(MAPAD_EINDEX);
    }
    {  // .tpi
      FrameReader r;
      if (!r.open(p + ".tpi")) return fail(MAPAD_EIO);
      uint64_t nc;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      bool ok = r.u64(nc);
      for (uint64_t i = 0; ok && i < nc; ++i) {
        uint64_t s, e, l;
        ok = r.u64(s) && r.u64(e) && r.u64(l);
        std::string name(ok ? l : 0, '\0');
        ok = ok && (l == 0 || r.read(&name[0], l));
        ix->contig_start.push_back(s); ix->contig_end.push_back(e); ix->contig_names.push_back(name);
      }
      r.close();
      if (!ok) return fail(MAPAD_EIO);
    }
    {  // .tos
      FrameReader r;
      if (!r.open(p + ".tos")) return fail(MAPAD_EIO);
      uint64_t no;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      bool ok = r.u64(no);
      for (uint64_t i = 0; ok && i < no; ++i) {
        uint64_t pos; uint8_t sym;
        ok = r.u64(pos) && r.u8(sym);
        ix->orig_pos.push_back(pos); ix->orig_sym.push_back(sym);
      }
      r.close();
      if (!ok) return fail(MAPAD_EIO);
    }
    {  // .tle and .trt are validated against what the BWT implies; .toc is not needed (ranks come from the BWT re-layout)
      FrameReader r;
      if (!r.open(p + ".tle")) return fail(MAPAD_EIO);
      uint64_t nl;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      bool ok = r.u64(nl) && nl <= 8;
      uint64_t less[8] = {0};
      for (uint64_t i = 0; ok && i < nl; ++i) ok = r.u64(less[i]);
      r.close();
      if (!ok) return fail(MAPAD_EIO);
      ix->derive_from_bwt();
      for (uint64_t i = 0; i < nl; ++i) if (less[i] != ix->less[i]) return fail(MAPAD_EINDEX);
    }
    {
      FrameReader r;
      if (!r.open(p + ".trt")) return fail(MAPAD_EIO);
      uint64_t nr;
      if (!read_version(r, rc)) { r.close(); return fail(rc); }
      bool ok = r.u64(nr) && nr <= 6;
      for (uint64_t i = 0; ok && i < nr; ++i) {
        uint64_t sym; uint8_t rank;
        ok = r.u64(sym) && r.u8(rank);
        static const char alpha[6] = {'$', 'A', 'C', 'G', 'T', 'X'};
        if (ok && (rank >= 6 || (uint8_t)alpha[rank] != sym)) { r.close(); return fail(MAPAD_EINDEX); }
      }
      r.close();
      if (!ok) return fail(MAPAD_EIO);
    }
  } catch (const std::bad_alloc&) {
    return fail(MAPAD_ENOMEM);
  }
  if (ix->sa_rate == 0 || ix->sa_sample.size() != (ix->n + ix->sa_rate - 1) / ix->sa_rate) return fail(MAPAD_EINDEX);
  ix->refresh_view();
  *out = reinterpret_cast<mapad_index*>(ix);
  return MAPAD_OK;
}

}  // extern "C"
