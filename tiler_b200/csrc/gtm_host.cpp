// gtm_host.cpp -- host-side verification I/O for the GTM stream: the command serialiser of
// TTilingEncoder.SaveStream (tilingencoder.pas:5177-5482), a decoder with the semantics of LoadStream (:4880-5175) /
// gtm.player.js:365-546, and an LZMA codec for the stream's chunk format (extern.pas:420-440: lc = 8, lp = 0, pb = 2,
// end marker, 13-byte "alone" header with the size field set to 0xFF..FF).  neither liblzma nor Python's lzma accepts
// lc = 8 (they require lc + lp <= 4), and the FreePascal host / node are not available here, hence this file.
//
// Plain C++ on the CPU, built into libtm_gtm.so.  It is NOT part of the GPU product path (libtm_gpu.so): in the
// reference these jobs stay in the FreePascal host (bitstream writer) and in the player.
#include "../../include/tm_gtm.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

extern "C" {

// ================================================================== LZMA range coder
namespace {

constexpr int kNumBitModelTotalBits = 11;
constexpr uint32_t kBitModelTotal = 1u << kNumBitModelTotalBits;
constexpr int kNumMoveBits = 5;
constexpr uint32_t kTopValue = 1u << 24;
constexpr int kNumStates = 12;
constexpr int kNumPosStatesMax = 16;
constexpr int kNumLenToPosStates = 4;
constexpr int kNumPosSlotBits = 6;
constexpr int kStartPosModelIndex = 4, kEndPosModelIndex = 14;
constexpr int kNumFullDistances = 1 << (kEndPosModelIndex >> 1);   // 128
constexpr int kNumAlignBits = 4;
constexpr int kMatchMinLen = 2;
constexpr int kMatchMaxLen = 273;
#ifndef TMH_LZMA_DEPTH
#define TMH_LZMA_DEPTH 16   // hash-chain candidates examined per position (48: 2 % smaller streams, 1.6x slower)
#endif
typedef uint16_t Prob;

struct LenProbs {
  Prob choice, choice2;
  Prob low[kNumPosStatesMax][8], mid[kNumPosStatesMax][8], high[256];
};

struct Model {
  int lc, lp, pb;
  std::vector<Prob> lit;
  Prob isMatch[kNumStates][kNumPosStatesMax], isRep[kNumStates], isRepG0[kNumStates], isRepG1[kNumStates], isRepG2[kNumStates],
      isRep0Long[kNumStates][kNumPosStatesMax];
  Prob posSlot[kNumLenToPosStates][1 << kNumPosSlotBits];
  Prob posSpecial[kNumFullDistances - kEndPosModelIndex + 1];
  Prob posAlign[1 << kNumAlignBits];
  LenProbs lenMatch, lenRep;
  void init(int lc_, int lp_, int pb_) {
    lc = lc_; lp = lp_; pb = pb_;
    lit.assign((size_t)0x300 << (lc + lp), kBitModelTotal / 2);
    auto fill = [](Prob *p, size_t n) { for (size_t i = 0; i < n; ++i) p[i] = kBitModelTotal / 2; };
    fill(&isMatch[0][0], sizeof isMatch / 2); fill(isRep, kNumStates); fill(isRepG0, kNumStates); fill(isRepG1, kNumStates);
    fill(isRepG2, kNumStates); fill(&isRep0Long[0][0], sizeof isRep0Long / 2); fill(&posSlot[0][0], sizeof posSlot / 2);
    fill(posSpecial, sizeof posSpecial / 2); fill(posAlign, sizeof posAlign / 2);
    fill(&lenMatch.choice, sizeof(LenProbs) / 2); fill(&lenRep.choice, sizeof(LenProbs) / 2);
  }
  Prob *litProbs(uint64_t pos, uint8_t prev) { return lit.data() + (size_t)0x300 * (((pos & ((1u << lp) - 1)) << lc) + (prev >> (8 - lc))); }
};

inline int stateAfterLit(int s) { return s < 4 ? 0 : (s < 10 ? s - 3 : s - 6); }
inline int stateAfterMatch(int s) { return s < 7 ? 7 : 10; }
inline int stateAfterRep(int s) { return s < 7 ? 8 : 11; }
inline int stateAfterShortRep(int s) { return s < 7 ? 9 : 11; }

// ------------------------------------------------------------------ encoder
struct RangeEnc {
  uint64_t low = 0; uint32_t range = 0xFFFFFFFFu; uint8_t cache = 0; uint64_t cacheSize = 1;
  uint8_t *out = nullptr, *out_end = nullptr;   // caller-provided buffer; `overflow` is set instead of writing past its end
  bool overflow = false;
  inline void put(uint8_t b) { if (out < out_end) *out++ = b; else overflow = true; }
  inline void shiftLow() {
    if ((uint32_t)low < 0xFF000000u || (uint32_t)(low >> 32) != 0) {
      uint8_t temp = cache;
      do { put((uint8_t)(temp + (uint8_t)(low >> 32))); temp = 0xFF; } while (--cacheSize != 0);
      cache = (uint8_t)((uint32_t)low >> 24);
    }
    cacheSize++;
    low = (uint64_t)((uint32_t)low << 8);
  }
  // one adaptive bit, branch-free on the bit value (the bits of image data are close to unpredictable for the CPU)
  inline void bit(Prob *p, uint32_t b) {
    const uint32_t pv = *p;
    const uint32_t bound = (range >> kNumBitModelTotalBits) * pv;
    const uint32_t mask = 0u - b;                              // 0 or 0xFFFFFFFF
    low += bound & mask;
    range = (bound & ~mask) | ((range - bound) & mask);
    // b == 0: p += (2048 - p) >> 5;  b == 1: p -= p >> 5
    const uint32_t up = (kBitModelTotal - pv) >> kNumMoveBits, down = pv >> kNumMoveBits;
    *p = (Prob)(pv + (up & ~mask) - (down & mask));
    if (range < kTopValue) { range <<= 8; shiftLow(); }       // one step always suffices: range >= 2^24 * 31 / 2^11 before it
  }
  void direct(uint32_t v, int nbits) {
    for (int i = nbits - 1; i >= 0; --i) {
      range >>= 1;
      if ((v >> i) & 1) low += range;
      if (range < kTopValue) { range <<= 8; shiftLow(); }
    }
  }
  void tree(Prob *probs, int nbits, uint32_t sym) {
    uint32_t m = 1;
    for (int i = nbits - 1; i >= 0; --i) { const uint32_t b = (sym >> i) & 1; bit(probs + m, b); m = (m << 1) | b; }
  }
  void treeRev(Prob *probs, int nbits, uint32_t sym) {
    uint32_t m = 1;
    for (int i = 0; i < nbits; ++i) { const uint32_t b = sym & 1; bit(probs + m, b); m = (m << 1) | b; sym >>= 1; }
  }
  void flush() { for (int i = 0; i < 5; ++i) shiftLow(); }
};

struct Encoder {
  Model m; RangeEnc rc; int state = 0; uint32_t reps[4] = {0, 0, 0, 0};
  const uint8_t *src; size_t n; uint32_t dictSize;

  void encLen(LenProbs &lp, uint32_t len, uint32_t posState) {
    len -= kMatchMinLen;
    if (len < 8) { rc.bit(&lp.choice, 0); rc.tree(lp.low[posState], 3, len); }
    else {
      rc.bit(&lp.choice, 1);
      if (len < 16) { rc.bit(&lp.choice2, 0); rc.tree(lp.mid[posState], 3, len - 8); }
      else { rc.bit(&lp.choice2, 1); rc.tree(lp.high, 8, len - 16); }
    }
  }
  void encLiteral(size_t pos) {
    const uint32_t posState = (uint32_t)pos & ((1u << m.pb) - 1);
    rc.bit(&m.isMatch[state][posState], 0);
    Prob *probs = m.litProbs(pos, pos ? src[pos - 1] : 0);
    uint32_t symbol = src[pos] | 0x100u;
    if (state < 7) {
      do { rc.bit(probs + (symbol >> 8), (symbol >> 7) & 1); symbol <<= 1; } while (symbol < 0x10000);
    } else {
      uint32_t matchByte = src[pos - reps[0] - 1], offs = 0x100;
      do {
        matchByte <<= 1;
        rc.bit(probs + (offs + (matchByte & offs) + (symbol >> 8)), (symbol >> 7) & 1);
        symbol <<= 1;
        offs &= ~(matchByte ^ symbol);
      } while (symbol < 0x10000);
    }
    state = stateAfterLit(state);
  }
  void encDistance(uint32_t dist, uint32_t len) {   // dist = distance - 1
    uint32_t slot;
    if (dist < 4) slot = dist;
    else { int nb = 31 - __builtin_clz(dist); slot = (uint32_t)(2 * nb) + ((dist >> (nb - 1)) & 1); }
    const uint32_t lenState = len - kMatchMinLen < kNumLenToPosStates - 1 ? len - kMatchMinLen : kNumLenToPosStates - 1;
    rc.tree(m.posSlot[lenState], kNumPosSlotBits, slot);
    if (slot >= kStartPosModelIndex) {
      const int footerBits = (int)(slot >> 1) - 1;
      const uint32_t base = (2 | (slot & 1)) << footerBits;
      const uint32_t reduced = dist - base;
      if (slot < kEndPosModelIndex) rc.treeRev(m.posSpecial + base - slot - 1, footerBits, reduced);
      else { rc.direct(reduced >> kNumAlignBits, footerBits - kNumAlignBits); rc.treeRev(m.posAlign, kNumAlignBits, reduced & 15); }
    }
  }
  void encMatch(size_t pos, uint32_t dist, uint32_t len) {
    const uint32_t posState = (uint32_t)pos & ((1u << m.pb) - 1);
    rc.bit(&m.isMatch[state][posState], 1);
    rc.bit(&m.isRep[state], 0);
    encLen(m.lenMatch, len, posState);
    encDistance(dist, len);
    reps[3] = reps[2]; reps[2] = reps[1]; reps[1] = reps[0]; reps[0] = dist;
    state = stateAfterMatch(state);
  }
  void encRep(size_t pos, int ri, uint32_t len) {
    const uint32_t posState = (uint32_t)pos & ((1u << m.pb) - 1);
    rc.bit(&m.isMatch[state][posState], 1);
    rc.bit(&m.isRep[state], 1);
    if (ri == 0) { rc.bit(&m.isRepG0[state], 0); rc.bit(&m.isRep0Long[state][posState], 1); }
    else {
      rc.bit(&m.isRepG0[state], 1);
      if (ri == 1) rc.bit(&m.isRepG1[state], 0);
      else { rc.bit(&m.isRepG1[state], 1); rc.bit(&m.isRepG2[state], ri - 2); }
      const uint32_t d = reps[ri];
      for (int i = ri; i > 0; --i) reps[i] = reps[i - 1];
      reps[0] = d;
    }
    encLen(m.lenRep, len, posState);
    state = stateAfterRep(state);
  }
  void encEndMarker(size_t pos) {
    const uint32_t posState = (uint32_t)pos & ((1u << m.pb) - 1);
    rc.bit(&m.isMatch[state][posState], 1);
    rc.bit(&m.isRep[state], 0);
    encLen(m.lenMatch, kMatchMinLen, posState);
    encDistance(0xFFFFFFFFu, kMatchMinLen);
  }
  uint32_t matchLen(size_t pos, size_t cand, uint32_t limit) const {
    uint32_t l = 0;
    while (l + 8 <= limit) {   // eight bytes at a time
      uint64_t x, y;
      memcpy(&x, src + cand + l, 8); memcpy(&y, src + pos + l, 8);
      if (x != y) return l + (uint32_t)(__builtin_ctzll(x ^ y) >> 3);
      l += 8;
    }
    while (l < limit && src[cand + l] == src[pos + l]) ++l;
    return l;
  }

  // ---- parse: greedy LZ77 over one BLOCK of the input, independent of every other block's parse.
  // The match finder is a hash chain over 4-byte prefixes built once for the whole input (prev[p] = the latest earlier
  // position with p's hash), so a block sees the complete history before it.  A block starts without repeat distances
  // (block 0: with the format's initial {0,0,0,0}) and no match crosses a block end; the decisions therefore depend on the
  // input and the fixed block size only -- never on the thread count -- and the stream is the same however many threads parse.
  struct Tok { uint32_t len, dist; };   // len == 1: literal; otherwise a match of distance dist + 1
  static constexpr size_t kBlock = (size_t)1 << 18;
  void parseBlock(size_t b0, size_t b1, const int32_t *prev, std::vector<Tok> &out) const {
    uint32_t lr[4] = {0, 0, 0, 0};
    int nrep = b0 == 0 ? 4 : 0;
    out.clear();
    out.reserve((b1 - b0) / 4 + 16);
    size_t pos = b0;
    while (pos < b1) {
      const uint32_t limit = (uint32_t)(b1 - pos < (size_t)kMatchMaxLen ? b1 - pos : (size_t)kMatchMaxLen);
      uint32_t bestLen = 0, bestDist = 0;
      if (pos > 0) {
        for (int r = 0; r < nrep; ++r) {
          if ((size_t)lr[r] + 1 > pos) continue;
          const uint32_t l = matchLen(pos, pos - lr[r] - 1, limit);
          if (l >= 2 && l > bestLen) { bestLen = l; bestDist = lr[r]; }
        }
      }
      if (pos + 4 <= n) {
        int64_t c = prev[pos];
        int depth = TMH_LZMA_DEPTH;
        uint32_t mlen = bestLen >= 3 ? bestLen : 3;   // a normal match must beat the best repeat (and be >= 4)
        while (c >= 0 && depth-- > 0) {
          const size_t d = pos - (size_t)c;
          if (d > dictSize) break;
          if (mlen >= limit) break;
          if (src[(size_t)c + mlen] == src[pos + mlen]) {
            const uint32_t l = matchLen(pos, (size_t)c, limit);
            if (l > mlen) { mlen = l; bestLen = l; bestDist = (uint32_t)(d - 1); if (l >= 128) break; }
          }
          c = prev[(size_t)c];
        }
      }
      if (bestLen < 2) { out.push_back({1u, 0u}); ++pos; continue; }
      out.push_back({bestLen, bestDist});
      int r = 0;
      while (r < nrep && lr[r] != bestDist) ++r;
      if (r == nrep) { if (nrep < 4) ++nrep; r = nrep - 1; }   // a new distance enters at the front, the oldest leaves
      for (int i = r; i > 0; --i) lr[i] = lr[i - 1];
      lr[0] = bestDist;
      pos += bestLen;
    }
  }

  // ---- code: the token stream of a block through the range coder.  A match whose distance is one of the coder's four
  // repeat distances is written as that repeat (the parser's own repeats always are: its list is a prefix of this one).
  void codeBlock(size_t b0, const std::vector<Tok> &toks) {
    size_t pos = b0;
    for (const Tok &t : toks) {
      if (t.len == 1) { encLiteral(pos); ++pos; continue; }
      int r = 0;
      while (r < 4 && !(reps[r] == t.dist && (size_t)reps[r] + 1 <= pos)) ++r;
      if (r < 4) encRep(pos, r, t.len); else encMatch(pos, t.dist, t.len);
      pos += t.len;
    }
  }

  void run(int n_threads) {
    constexpr int HB = 20;
    std::vector<int32_t> head((size_t)1 << HB, -1), prev(n ? n : 1, -1);
    auto h4 = [&](size_t p) { uint32_t v; memcpy(&v, src + p, 4); return (v * 2654435761u) >> (32 - HB); };
    for (size_t p = 0; p + 4 <= n; ++p) { const uint32_t h = h4(p); prev[p] = head[h]; head[h] = (int32_t)p; }
    const size_t nb = (n + kBlock - 1) / kBlock;
    std::vector<std::vector<Tok>> toks(nb);
    if (n_threads > (int)nb) n_threads = (int)nb;
    if (n_threads <= 1) {
      for (size_t b = 0; b < nb; ++b) {
        parseBlock(b * kBlock, std::min(n, (b + 1) * kBlock), prev.data(), toks[b]);
        codeBlock(b * kBlock, toks[b]);
        std::vector<Tok>().swap(toks[b]);
      }
    } else {
      // parser threads take blocks in order; this thread codes block b as soon as it is parsed
      std::atomic<size_t> next{0};
      std::vector<std::atomic<int>> done(nb);
      for (auto &d : done) d.store(0, std::memory_order_relaxed);
      std::mutex mu; std::condition_variable cv;
      std::vector<std::thread> pool;
      for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
          for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nb) break;
            parseBlock(b * kBlock, std::min(n, (b + 1) * kBlock), prev.data(), toks[b]);
            { std::lock_guard<std::mutex> lk(mu); done[b].store(1, std::memory_order_release); }
            cv.notify_all();
          }
        });
      for (size_t b = 0; b < nb; ++b) {
        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done[b].load(std::memory_order_acquire) != 0; }); }
        codeBlock(b * kBlock, toks[b]);
        std::vector<Tok>().swap(toks[b]);
      }
      for (auto &th : pool) th.join();
    }
    encEndMarker(n);
    rc.flush();
  }
};

// ------------------------------------------------------------------ decoder
struct RangeDec {
  const uint8_t *p, *end; uint32_t range = 0xFFFFFFFFu, code = 0; bool err = false;
  uint8_t next() { if (p >= end) { err = true; return 0; } return *p++; }
  void init() { next(); for (int i = 0; i < 4; ++i) code = (code << 8) | next(); }
  void norm() { if (range < kTopValue) { range <<= 8; code = (code << 8) | next(); } }
  uint32_t bit(Prob *pr) {
    const uint32_t bound = (range >> kNumBitModelTotalBits) * *pr;
    uint32_t b;
    if (code < bound) { range = bound; *pr = (Prob)(*pr + ((kBitModelTotal - *pr) >> kNumMoveBits)); b = 0; }
    else { range -= bound; code -= bound; *pr = (Prob)(*pr - (*pr >> kNumMoveBits)); b = 1; }
    norm();
    return b;
  }
  uint32_t direct(int nbits) {
    uint32_t r = 0;
    for (; nbits > 0; --nbits) {
      range >>= 1; code -= range;
      const uint32_t t = 0u - (code >> 31);
      code += range & t;
      r = (r << 1) + (t + 1);
      norm();
    }
    return r;
  }
  uint32_t tree(Prob *probs, int nbits) { uint32_t m = 1; for (int i = 0; i < nbits; ++i) m = (m << 1) | bit(probs + m); return m - (1u << nbits); }
  uint32_t treeRev(Prob *probs, int nbits) {
    uint32_t m = 1, s = 0;
    for (int i = 0; i < nbits; ++i) { const uint32_t b = bit(probs + m); m = (m << 1) | b; s |= b << i; }
    return s;
  }
};

uint32_t decLen(RangeDec &rc, LenProbs &lp, uint32_t posState) {
  if (rc.bit(&lp.choice) == 0) return rc.tree(lp.low[posState], 3);
  if (rc.bit(&lp.choice2) == 0) return 8 + rc.tree(lp.mid[posState], 3);
  return 16 + rc.tree(lp.high, 8);
}

}  // namespace

// src -> LZMA "alone" stream (props byte, dict size, 8 x 0xFF, range-coded data with end marker).  Returns the size, or -1
// when out_cap is too small (call again with a larger buffer; worst case is about n + n/8 + 64).
int64_t tmh_lzma_encode(const uint8_t *src, int64_t n, int lc, int lp, int pb, uint32_t dict_size, uint8_t *out, int64_t out_cap) {
  return tmh_lzma_encode_mt(src, n, lc, lp, pb, dict_size, out, out_cap, 0);
}
int64_t tmh_lzma_encode_mt(const uint8_t *src, int64_t n, int lc, int lp, int pb, uint32_t dict_size, uint8_t *out, int64_t out_cap,
                           int n_threads) {
  if (n < 0 || lc < 0 || lc > 8 || lp < 0 || lp > 4 || pb < 0 || pb > 4) return -2;
  if (n > 0x7fff0000 || out_cap < 13 + 8) return n > 0x7fff0000 ? -2 : -1;   // 32-bit chain links; header + flush always fit
  uint8_t *o = out;
  *o++ = (uint8_t)((pb * 5 + lp) * 9 + lc);
  for (int i = 0; i < 4; ++i) *o++ = (uint8_t)(dict_size >> (8 * i));
  for (int i = 0; i < 8; ++i) *o++ = 0xFF;
  Encoder e;
  e.m.init(lc, lp, pb);
  e.rc.out = o; e.rc.out_end = out + out_cap;
  e.src = src; e.n = (size_t)n; e.dictSize = dict_size;
  // parser threads for this stream (TMH_LZMA_THREADS overrides): up to 8, never more than the cores present
  static const int env_threads = getenv("TMH_LZMA_THREADS") ? atoi(getenv("TMH_LZMA_THREADS")) : 0;
  int nt = n_threads > 0 ? n_threads
                         : (env_threads > 0 ? env_threads : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())));
  e.run(nt);
  if (e.rc.overflow) return -1;
  return (int64_t)(e.rc.out - out);
}

// Decodes ONE "alone" stream starting at src (stops at the end marker, or at the header's size when it is known).
// Returns the decoded size (-1: output too small, -2: corrupt input); *consumed = input bytes used, so that the
// back-to-back streams of a GTM file can be walked (wlzma.wrk.js:50-63).
int64_t tmh_lzma_decode(const uint8_t *src, int64_t n, uint8_t *out, int64_t out_cap, int64_t *consumed) {
  if (n < 13 + 5) return -2;
  int d = src[0];
  if (d >= 9 * 5 * 5) return -2;
  const int lc = d % 9; d /= 9;
  const int lp = d % 5, pb = d / 5;
  uint64_t unpack = 0;
  for (int i = 0; i < 8; ++i) unpack |= (uint64_t)src[5 + i] << (8 * i);
  const bool known = unpack != ~0ull;
  Model m;
  m.init(lc, lp, pb);
  RangeDec rc;
  rc.p = src + 13; rc.end = src + n;
  rc.init();
  int state = 0;
  uint32_t rep0 = 0, rep1 = 0, rep2 = 0, rep3 = 0;
  int64_t pos = 0;
  for (;;) {
    if (known && (uint64_t)pos >= unpack) break;
    if (rc.err) return -2;
    const uint32_t posState = (uint32_t)pos & ((1u << pb) - 1);
    if (rc.bit(&m.isMatch[state][posState]) == 0) {
      if (pos >= out_cap) return -1;
      Prob *probs = m.litProbs((uint64_t)pos, pos ? out[pos - 1] : 0);
      uint32_t symbol = 1;
      if (state >= 7) {
        uint32_t matchByte = out[pos - rep0 - 1];
        do {
          const uint32_t mb = (matchByte >> 7) & 1;
          matchByte <<= 1;
          const uint32_t b = rc.bit(probs + (((1 + mb) << 8) + symbol));
          symbol = (symbol << 1) | b;
          if (mb != b) break;
        } while (symbol < 0x100);
      }
      while (symbol < 0x100) symbol = (symbol << 1) | rc.bit(probs + symbol);
      out[pos++] = (uint8_t)symbol;
      state = stateAfterLit(state);
      continue;
    }
    uint32_t len;
    if (rc.bit(&m.isRep[state]) != 0) {
      if (pos == 0) return -2;
      if (rc.bit(&m.isRepG0[state]) == 0) {
        if (rc.bit(&m.isRep0Long[state][posState]) == 0) {
          if (pos >= out_cap) return -1;
          state = stateAfterShortRep(state);
          out[pos] = out[pos - rep0 - 1];
          ++pos;
          continue;
        }
      } else {
        uint32_t dist;
        if (rc.bit(&m.isRepG1[state]) == 0) dist = rep1;
        else {
          if (rc.bit(&m.isRepG2[state]) == 0) dist = rep2;
          else { dist = rep3; rep3 = rep2; }
          rep2 = rep1;
        }
        rep1 = rep0;
        rep0 = dist;
      }
      len = decLen(rc, m.lenRep, posState);
      state = stateAfterRep(state);
    } else {
      rep3 = rep2; rep2 = rep1; rep1 = rep0;
      len = decLen(rc, m.lenMatch, posState);
      state = stateAfterMatch(state);
      const uint32_t lenState = len < kNumLenToPosStates - 1 ? len : kNumLenToPosStates - 1;
      const uint32_t slot = rc.tree(m.posSlot[lenState], kNumPosSlotBits);
      if (slot < kStartPosModelIndex) rep0 = slot;
      else {
        const int footerBits = (int)(slot >> 1) - 1;
        rep0 = (2 | (slot & 1)) << footerBits;
        if (slot < kEndPosModelIndex) rep0 += rc.treeRev(m.posSpecial + rep0 - slot - 1, footerBits);
        else { rep0 += rc.direct(footerBits - kNumAlignBits) << kNumAlignBits; rep0 += rc.treeRev(m.posAlign, kNumAlignBits); }
      }
      if (rep0 == 0xFFFFFFFFu) break;   // end marker
    }
    len += kMatchMinLen;
    if ((int64_t)rep0 >= pos) return -2;
    if (pos + (int64_t)len > out_cap) return -1;
    for (uint32_t i = 0; i < len; ++i, ++pos) out[pos] = out[pos - rep0 - 1];
  }
  if (rc.err) return -2;
  if (consumed) *consumed = (int64_t)(rc.p - src);
  return pos;
}

// ================================================================== GTM commands (tilingencoder.pas:53-86, 5200-5268, 5394-5443)
enum { gtPredictedTileShortOffsets = 0, gtPredictedTileLongOffsets = 1, gtShortTileIdxShortPalIdx = 2, gtLongTileIdxShortPalIdx = 3,
       gtLongTileIdxLongPalIdx = 4, gtIntraTile = 5, gtSkipBlock = 6, gtFrameEnd = 11, gtLoadPalette = 12, gtTileSet = 13,
       gtSetDimensions = 14, gtExtendedCommand = 15 };

namespace {
struct W {
  uint8_t *p; int64_t cap, n = 0;
  void b(uint8_t v) { if (n < cap) p[n] = v; ++n; }
  void w(uint32_t v) { b((uint8_t)v); b((uint8_t)(v >> 8)); }
  void d(uint32_t v) { w(v & 0xffff); w(v >> 16); }
  void cmd(int c, uint32_t data) { w((data << 4) | (uint32_t)c); }
};
}  // namespace

// The tilemap items of n_frames consecutive frames of one keyframe sequence -> command bytes (DoTMI, SkipBlock runs,
// FrameEnd; the last frame gets the keyframe-end bit when last_is_kf_end != 0).  tiles [n_tiles][64] palette indices and
// use_count [n_tiles] describe the final (re-indexed) dictionary: a tile used once is sent inline as IntraTile (:5236).
// mirror bit 0 = HMirror, bit 1 = VMirror.  Returns the byte count (> cap: buffer too small, nothing is lost but the tail).
int64_t tmh_gtm_write_frames(const int32_t *tile_idx, const int32_t *pal_idx, const int32_t *pred_x, const int32_t *pred_y,
                             const uint8_t *is_pred, const uint8_t *mirror, int n_frames, int tiles_per_frame, const uint8_t *tiles,
                             const int32_t *use_count, int64_t n_tiles, int emit_skip_blocks, int last_is_kf_end, uint8_t *out,
                             int64_t cap) {
  W o{out, cap};
  const int kMinSkip = 4, kMaxSkip = 1 << 12;
  for (int f = 0; f < n_frames; ++f) {
    const int64_t base = (int64_t)f * tiles_per_frame;
    int skip = 0;
    for (int yx = 0; yx < tiles_per_frame; ++yx) {
      if (skip > 0) { --skip; continue; }
      int run = 0;
      if (emit_skip_blocks)
        for (int s = yx; s < tiles_per_frame; ++s) {
          const int64_t i = base + s;
          if (!(is_pred[i] && pred_x[i] == 0 && pred_y[i] == 0)) break;   // IsSmoothed, :621-624
          ++run;
        }
      if (run > kMaxSkip) run = kMaxSkip;
      if (run >= kMinSkip) { o.cmd(gtSkipBlock, (uint32_t)(run - 1)); skip = run - 1; continue; }
      const int64_t i = base + yx;
      if (is_pred[i]) {
        const int px = pred_x[i], py = pred_y[i];
        if (px < -32 || px > 31 || py < -32 || py > 31) { o.cmd(gtPredictedTileLongOffsets, 0); o.b((uint8_t)(int8_t)px); o.b((uint8_t)(int8_t)py); }
        else o.cmd(gtPredictedTileShortOffsets, ((uint32_t)px & 63) | (((uint32_t)py & 63) << 6));
      } else {
        const uint32_t ti = tile_idx[i] > 0 ? (uint32_t)tile_idx[i] : 0, pi = pal_idx[i] > 0 ? (uint32_t)pal_idx[i] : 0;
        const bool intra = (int64_t)ti < n_tiles && use_count[ti] <= 1;
        const uint32_t attrs = mirror[i] & 3;   // (VMirror << 1) | HMirror
        if (intra) { o.cmd(gtIntraTile, attrs); o.w(pi); for (int k = 0; k < 64; ++k) o.b(tiles[(int64_t)ti * 64 + k]); }
        else if (ti <= 0xffff && pi < (1u << 10)) { o.cmd(gtShortTileIdxShortPalIdx, attrs | (pi << 2)); o.w(ti); }
        else if (pi < (1u << 10)) { o.cmd(gtLongTileIdxShortPalIdx, attrs | (pi << 2)); o.d(ti); }
        else { o.cmd(gtLongTileIdxLongPalIdx, attrs); o.w(pi); o.d(ti); }
      }
    }
    o.cmd(gtFrameEnd, (f == n_frames - 1 && last_is_kf_end) ? 1u : 0u);
  }
  return o.n;
}

// Decoder state + one call that plays a raw (decompressed) command buffer.  Semantics of gtm.player.js:365-546 and
// LoadStream (:4880-5175): mirrored copies are resolved at draw time, predicted tiles copy from the previous output frame,
// IntraTile tiles live in a ring of 2*W*H slots after the dictionary.  Frames are packed 0x00BBGGRR like the encoder's.
struct tmh_gtm_decoder {
  int w = 0, h = 0; uint32_t tile_count = 0, total_tiles = 0, cur_intra = 0; int pal_size = 0;
  std::vector<uint8_t> tiles; std::vector<std::vector<int32_t>> pals;
  std::vector<int32_t> buf[2]; int cur = 0; int tm_pos = 0; int64_t frames = 0;
};
tmh_gtm_decoder *tmh_gtm_decoder_create(void) { return new tmh_gtm_decoder(); }
void tmh_gtm_decoder_destroy(tmh_gtm_decoder *d) { delete d; }
int tmh_gtm_decoder_dims(tmh_gtm_decoder *d, int *w, int *h, int64_t *frames) { *w = d->w; *h = d->h; *frames = d->frames; return 0; }

static void draw_tile(tmh_gtm_decoder *d, uint32_t idx, uint32_t attrs) {
  const uint32_t pal = attrs >> 2;
  if (idx >= d->total_tiles || pal >= d->pals.size() || d->pals[pal].empty()) { d->tm_pos++; return; }
  const uint8_t *t = d->tiles.data() + (size_t)idx * 64;
  const int32_t *p = d->pals[pal].data();
  const int x0 = (d->tm_pos % d->w) * 8, y0 = (d->tm_pos / d->w) * 8, W8 = d->w * 8;
  int32_t *dst = d->buf[d->cur].data();
  for (int ty = 0; ty < 8; ++ty)
    for (int tx = 0; tx < 8; ++tx) {
      const int sx = (attrs & 1) ? 7 - tx : tx, sy = (attrs & 2) ? 7 - ty : ty;
      const uint8_t v = t[sy * 8 + sx];
      dst[(size_t)(y0 + ty) * W8 + x0 + tx] = v < d->pal_size ? p[v] : 0;
    }
  d->tm_pos++;
}
static void draw_pred(tmh_gtm_decoder *d, int ox, int oy) {
  const int x0 = (d->tm_pos % d->w) * 8, y0 = (d->tm_pos / d->w) * 8, W8 = d->w * 8, H8 = d->h * 8;
  int32_t *dst = d->buf[d->cur].data();
  const int32_t *src = d->buf[1 - d->cur].data();
  for (int ty = 0; ty < 8; ++ty)
    for (int tx = 0; tx < 8; ++tx) {
      const int sy = y0 + ty + oy, sx = x0 + tx + ox;
      dst[(size_t)(y0 + ty) * W8 + x0 + tx] = (sy >= 0 && sy < H8 && sx >= 0 && sx < W8) ? src[(size_t)sy * W8 + sx] : 0;
    }
  d->tm_pos++;
}

// Plays `n` command bytes; every completed frame is appended to frames_out (capacity max_frames frames of w*8 x h*8).
// Returns the number of frames produced by this call, or -1 on a malformed stream.
int64_t tmh_gtm_decode(tmh_gtm_decoder *d, const uint8_t *s, int64_t n, int32_t *frames_out, int64_t max_frames) {
  int64_t p = 0, produced = 0;
  auto rb = [&]() -> uint32_t { return p < n ? s[p++] : (p++, 0u); };
  auto rw = [&]() -> uint32_t { uint32_t v = rb(); v |= rb() << 8; return v; };
  auto rd = [&]() -> uint32_t { uint32_t v = rw(); v |= rw() << 16; return v; };
  while (p < n) {
    const uint32_t v = rw();
    const uint32_t c = v & 15, data = v >> 4;
    switch (c) {
      case gtSetDimensions: {
        d->w = (int)rw(); d->h = (int)rw(); rd();
        d->tile_count = rd();
        d->cur_intra = d->tile_count;
        d->total_tiles = d->tile_count + (uint32_t)(d->w * d->h * 2);
        d->tiles.assign((size_t)d->total_tiles * 64, 0);
        d->buf[0].assign((size_t)d->w * d->h * 64, 0); d->buf[1].assign((size_t)d->w * d->h * 64, 0);
        break;
      }
      case gtTileSet: {
        const uint32_t ts = rd(), te = rd();
        d->pal_size = (int)data;
        for (uint32_t t = ts; t <= te && p < n; ++t)
          for (int k = 0; k < 64; ++k) { const uint8_t b = (uint8_t)rb(); if (t < d->total_tiles) d->tiles[(size_t)t * 64 + k] = b; }
        break;
      }
      case gtLoadPalette: {
        const uint32_t pi = rw();
        if (pi >= d->pals.size()) d->pals.resize(pi + 1);
        d->pals[pi].assign((size_t)d->pal_size, 0);
        for (int i = 0; i < d->pal_size; ++i) { const uint32_t r = rb(), g = rb(), b = rb(); rb(); d->pals[pi][i] = (int32_t)(r | (g << 8) | (b << 16)); }
        break;
      }
      case gtFrameEnd: {
        if (d->w == 0 || d->tm_pos != d->w * d->h) return -1;
        if (produced < max_frames && frames_out) memcpy(frames_out + (size_t)produced * d->buf[0].size(), d->buf[d->cur].data(), d->buf[0].size() * 4);
        ++produced; ++d->frames;
        d->tm_pos = 0; d->cur = 1 - d->cur;
        break;
      }
      case gtSkipBlock: {
        if (d->w == 0) return -1;
        for (uint32_t i = 0; i < data + 1 && d->tm_pos < d->w * d->h; ++i) draw_pred(d, 0, 0);
        break;
      }
      case gtShortTileIdxShortPalIdx: if (d->w == 0) return -1; draw_tile(d, rw(), data); break;
      case gtLongTileIdxShortPalIdx: if (d->w == 0) return -1; draw_tile(d, rd(), data); break;
      case gtLongTileIdxLongPalIdx: { if (d->w == 0) return -1; const uint32_t pw = rw(); draw_tile(d, rd(), data | (pw << 2)); break; }
      case gtPredictedTileShortOffsets:
        if (d->w == 0) return -1;
        draw_pred(d, (int)(data & 31) - (int)(data & 32), (int)((data >> 6) & 31) - (int)((data >> 6) & 32));
        break;
      case gtPredictedTileLongOffsets: { if (d->w == 0) return -1; const int ox = (int8_t)rb(), oy = (int8_t)rb(); draw_pred(d, ox, oy); break; }
      case gtIntraTile: {
        if (d->w == 0) return -1;
        const uint32_t pw = rw();
        for (int k = 0; k < 64; ++k) d->tiles[(size_t)d->cur_intra * 64 + k] = (uint8_t)rb();
        draw_tile(d, d->cur_intra, data | (pw << 2));
        if (++d->cur_intra >= d->total_tiles) d->cur_intra = d->total_tiles - (uint32_t)(d->w * d->h * 2);
        break;
      }
      case gtExtendedCommand: { const uint32_t sz = rd(); p += sz; break; }
      default: return -1;
    }
  }
  return p == n ? produced : -1;
}

}  // extern "C"
