// dev_index_build.cpp — K0 of SURVEY §2: one-off re-layout of BWT + Occ + sampled SA
// (/root/reference/src/index/indexing.rs:163-210 produce them as separate arrays) into
// sector-sized interleaved count + 2-bit-code blocks (common.h).
#include "dev_index_build.hpp"

#include <cstring>

#include "../../include/mapad_gpu.h"

namespace mapad {

static uint64_t align64(uint64_t x) { return (x + 63) & ~63ull; }

int build_device_blob(const HostIndex& ix, IndexMeta& m, std::vector<uint8_t>& blob, int layout) {
  memset(&m, 0, sizeof m);
  const uint64_t n = ix.n;
  if (n < 2 || ix.sa_rate == 0 || ix.sa_rate > 0xffffffffull) return MAPAD_EINDEX;
  m.n = n;
  for (int i = 0; i < 8; ++i) m.less[i] = ix.less[i];
  m.sentinel_rows[0] = ix.sentinel_rows[0];
  m.sentinel_rows[1] = ix.sentinel_rows[1];
  m.wide = n >= (1ull << 31) ? 1u : 0u;
  if (layout == 1) m.wide = 1u;
  if (layout == 0 && m.wide) return MAPAD_EINDEX;
  m.sa_rate = (uint32_t)ix.sa_rate;
  const uint32_t syms = m.wide ? 128 : 64;
  const uint32_t bsz = m.wide ? 64 : 32;
  m.n_blocks = (n + syms - 1) / syms;
  m.n_sa = ix.sa_sample.size();
  m.n_extra = ix.extra_rows.size() / 2;
  m.n_contigs = ix.contig_start.size();
  m.n_orig = ix.orig_pos.size();
  // X ranges
  std::vector<XRange> xr;
  {
    uint64_t before = 0, i = 0;
    while (i < n) {
      if (ix.bwt[i] == 5) {
        uint64_t j = i;
        while (j < n && ix.bwt[j] == 5) ++j;
        xr.push_back(XRange{i, j, before});
        before += j - i;
        i = j;
      } else {
        ++i;
      }
    }
    m.n_x_rows = before;
  }
  m.n_xranges = xr.size();
  uint64_t off = 0;
  m.off_occ = off; off = align64(off + m.n_blocks * bsz);
  m.off_sa = off; off = align64(off + m.n_sa * (m.wide ? 8 : 4));
  m.off_extra = off; off = align64(off + m.n_extra * 16);
  m.off_xranges = off; off = align64(off + m.n_xranges * sizeof(XRange));
  m.off_contigs = off; off = align64(off + m.n_contigs * 16);
  m.off_orig = off; off = align64(off + m.n_orig * 9);
  m.total_bytes = off;
  blob.assign(off, 0);
  // occ blocks
  uint64_t cnt[4] = {0, 0, 0, 0};
  for (uint64_t b = 0; b < m.n_blocks; ++b) {
    uint8_t* p = blob.data() + m.off_occ + b * bsz;
    uint32_t* codes;
    if (m.wide) {
      uint64_t* c = reinterpret_cast<uint64_t*>(p);
      for (int k = 0; k < 4; ++k) c[k] = cnt[k];
      codes = reinterpret_cast<uint32_t*>(p + 32);
    } else {
      uint32_t* c = reinterpret_cast<uint32_t*>(p);
      for (int k = 0; k < 4; ++k) c[k] = (uint32_t)cnt[k];
      codes = reinterpret_cast<uint32_t*>(p + 16);
    }
    bool has_x = false;
    const uint64_t lo = b * syms, hi = lo + syms < n ? lo + syms : n;
    for (uint64_t i = lo; i < hi; ++i) {
      uint8_t s = ix.bwt[i];
      uint32_t code = 0;
      if (s >= 1 && s <= 4) { code = s - 1; cnt[s - 1] += 1; }
      else if (s == 5) has_x = true;
      else if (s != 0) return MAPAD_EINDEX;
      uint32_t w = (uint32_t)(i - lo);
      codes[w >> 4] |= code << (2 * (w & 15));
    }
    if (has_x) {
      if (m.wide) reinterpret_cast<uint64_t*>(p)[0] |= 1ull << 63;
      else reinterpret_cast<uint32_t*>(p)[0] |= 1u << 31;
    }
  }
  // sampled SA
  if (m.wide) memcpy(blob.data() + m.off_sa, ix.sa_sample.data(), m.n_sa * 8);
  else {
    uint32_t* s = reinterpret_cast<uint32_t*>(blob.data() + m.off_sa);
    for (uint64_t i = 0; i < m.n_sa; ++i) s[i] = (uint32_t)ix.sa_sample[i];
  }
  if (m.n_extra) memcpy(blob.data() + m.off_extra, ix.extra_rows.data(), m.n_extra * 16);
  if (m.n_xranges) memcpy(blob.data() + m.off_xranges, xr.data(), m.n_xranges * sizeof(XRange));
  {
    uint64_t* c = reinterpret_cast<uint64_t*>(blob.data() + m.off_contigs);
    for (uint64_t i = 0; i < m.n_contigs; ++i) { c[2 * i] = ix.contig_start[i]; c[2 * i + 1] = ix.contig_end[i]; }
  }
  if (m.n_orig) {
    memcpy(blob.data() + m.off_orig, ix.orig_pos.data(), m.n_orig * 8);
    memcpy(blob.data() + m.off_orig + 8 * m.n_orig, ix.orig_sym.data(), m.n_orig);
  }
  return MAPAD_OK;
}

}  // namespace mapad
