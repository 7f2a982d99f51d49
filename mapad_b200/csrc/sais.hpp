// sais.hpp — suffix array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009), written
// for this project.  Integer alphabet, recursion on reduced strings, O(n) time.
// Replaces rust-bio's `suffix_array()` call of the reference indexer
// (/root/reference/src/index/indexing.rs:163); symbol order with two sentinels follows rust-bio's
// transform_text (last '$' smallest, first '$' second smallest) and is applied by the caller.
#pragma once
#include <cstdint>
#include <vector>

namespace mapad {

template <class Int, class Sym>
class Sais {
 public:
  // s[0..n) over alphabet [0,K); s[n-1] must be the unique smallest symbol.  sa must hold n entries.
  static void build(const Sym* s, Int* sa, Int n, Int K) {
    if (n == 0) return;
    if (n == 1) { sa[0] = 0; return; }
    std::vector<uint64_t> tbits(((uint64_t)n + 63) / 64, 0);  // 1 = S-type
    auto is_s = [&](Int i) -> bool { return (tbits[(uint64_t)i >> 6] >> ((uint64_t)i & 63)) & 1; };
    auto set_s = [&](Int i) { tbits[(uint64_t)i >> 6] |= 1ull << ((uint64_t)i & 63); };
    set_s(n - 1);
    for (Int i = n - 2; i >= 0; --i) {
      if (s[i] < s[i + 1] || (s[i] == s[i + 1] && is_s(i + 1))) set_s(i);
    }
    auto is_lms = [&](Int i) -> bool { return i > 0 && is_s(i) && !is_s(i - 1); };

    std::vector<Int> bkt((size_t)K);
    auto buckets = [&](bool end) {
      for (Int c = 0; c < K; ++c) bkt[(size_t)c] = 0;
      for (Int i = 0; i < n; ++i) bkt[(size_t)s[i]]++;
      Int sum = 0;
      for (Int c = 0; c < K; ++c) {
        sum += bkt[(size_t)c];
        bkt[(size_t)c] = end ? sum : sum - bkt[(size_t)c];
      }
    };
    auto induce = [&]() {
      buckets(false);
      for (Int i = 0; i < n; ++i) {
        Int j = sa[i] - 1;
        if (sa[i] > 0 && !is_s(j)) sa[bkt[(size_t)s[j]]++] = j;
      }
      buckets(true);
      for (Int i = n - 1; i >= 0; --i) {
        Int j = sa[i] - 1;
        if (sa[i] > 0 && is_s(j)) sa[--bkt[(size_t)s[j]]] = j;
      }
    };

    // stage 1: sort LMS substrings
    for (Int i = 0; i < n; ++i) sa[i] = -1;
    buckets(true);
    for (Int i = 1; i < n; ++i)
      if (is_lms(i)) sa[--bkt[(size_t)s[i]]] = i;
    induce();
    // compact sorted LMS substrings into sa[0..n1)
    Int n1 = 0;
    for (Int i = 0; i < n; ++i)
      if (sa[i] >= 0 && is_lms(sa[i])) sa[n1++] = sa[i];
    for (Int i = n1; i < n; ++i) sa[i] = -1;
    // name them
    Int name = 0, prev = -1;
    for (Int i = 0; i < n1; ++i) {
      Int pos = sa[i];
      bool diff = false;
      if (prev < 0) diff = true;
      else {
        for (Int d = 0;; ++d) {
          if (pos + d >= n || prev + d >= n || s[pos + d] != s[prev + d] || is_s(pos + d) != is_s(prev + d)) { diff = true; break; }
          if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) {
            diff = !(is_lms(pos + d) && is_lms(prev + d));
            break;
          }
        }
      }
      if (diff) { ++name; prev = pos; }
      sa[n1 + pos / 2] = name - 1;
    }
    // gather reduced string into the tail sa[n-n1..n)
    for (Int i = n - 1, j = n - 1; i >= n1; --i)
      if (sa[i] >= 0) sa[j--] = sa[i];
    Int* s1 = sa + (n - n1);
    Int* sa1 = sa;
    // stage 2: solve the reduced problem
    if (name < n1) {
      Sais<Int, Int>::build(s1, sa1, n1, name);
    } else {
      for (Int i = 0; i < n1; ++i) sa1[s1[i]] = i;
    }
    // stage 3: induce the final order
    buckets(true);
    for (Int i = 1, j = 0; i < n; ++i)
      if (is_lms(i)) s1[j++] = i;
    for (Int i = 0; i < n1; ++i) sa1[i] = s1[sa1[i]];
    for (Int i = n1; i < n; ++i) sa[i] = -1;
    for (Int i = n1 - 1; i >= 0; --i) {
      Int j = sa[i];
      sa[i] = -1;
      sa[--bkt[(size_t)s[j]]] = j;
    }
    induce();
  }
};

}  // namespace mapad
