// knn_i8.cu -- exact k-nearest dictionary tiles under the reference's uint32 squared-Euclidean distance over
// 192 int16 coefficients (utils.pas:541-557; the ANN_short.dll contract, extern.pas:182-185), on sm_100a
// tensor cores.
//
//   d(q, t) = |q|^2 + |t|^2 - 2 q.t   (mod 2^32, exactly like the reference's Cardinal accumulator)
//
// Exactness on tensor cores: every int16 v is split into two 8-bit limbs v = 256*hi + lo (hi signed, lo unsigned).
// q.t = 65536*HH + 256*(HL + LH) + LL with HH, HL+LH, LL accumulated in int32 by tcgen05.mma kind::i8 -- no
// rounding anywhere.  A vector is stored as one 384-byte row [hi(192) | lo(192)] = 3 TMA/UMMA 128-byte swizzle
// atoms, so any K-step of A can be paired with any K-step of B by descriptor arithmetic alone.
//
// CTA (persistent, one per SM) = 128 query rows x the whole dictionary.  The query limb rows live in TMEM (tcgen05.st) and are
// the A operand of TS-mode MMAs; the dictionary streams through shared memory in 64-row tiles (24 KB) on an 8-stage TMA
// ring, 24 UTCIMMA (M128 N64 K32) per tile into one of two TMEM accumulator stages (3 x 64 columns: HH, X = HL + LH, LL).
// Eight epilogue warps (two per TMEM lane quarter, 32 of a tile's 64 columns each) fold the accumulators into distances and
// keep, per (query row, column half), either a running best list in registers (k = 1, k = 4) or a thresholded candidate
// strip in an L2-resident workspace that is cut back to the k best by scheduled, row-wide loose selections (2 <= k <= 64).
#include "tc_common.cuh"
#include "tm_kernels.h"
#include <cstdlib>
#include <cstdio>

namespace tmg {

constexpr int BM = 128;          // rows per M tile (UMMA M)
constexpr int BN = 64;           // rows per N tile (UMMA N)
constexpr int STAGES_K1 = 8;     // dictionary ring depth (k = 1 path)
constexpr int ROWB = 384;        // bytes per limb row
constexpr int CHUNK_A = BM * 128;  // one 128-byte-wide swizzle chunk of an A tile
constexpr int CHUNK_B = BN * 128;
constexpr int A_TILE = 3 * CHUNK_A;  // 49152
constexpr int B_TILE = 3 * CHUNK_B;  // 24576
constexpr int ACC_COLS = 3 * BN; // TMEM columns per accumulator stage
constexpr int A_COL = 2 * ACC_COLS;  // query rows live in TMEM columns [384, 480)

// ------------------------------------------------------------------ limb split + norms
// in: [n][192] int16 -> limbs [n][384] (hi bytes then lo bytes), norms[n] = sum v^2 mod 2^32
__global__ void __launch_bounds__(192) limb_split_kernel(const int16_t *__restrict__ in, int64_t n, uint8_t *__restrict__ limbs,
                                                        uint32_t *__restrict__ norms, uint32_t *__restrict__ norm_max) {
  __shared__ unsigned long long s_norm[8];   // exact (64-bit) squared norms: 192 * 32768^2 > 2^32
  const int r = threadIdx.x / 24, seg = threadIdx.x % 24;
  const int64_t row = (int64_t)blockIdx.x * 8 + r;
  if (threadIdx.x < 8) s_norm[threadIdx.x] = 0;
  __syncthreads();
  if (row < n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + row * 192) + seg);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[2], lo[2];
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t a = w[2 * i], b = w[2 * i + 1];
      // bytes of a: [lo0 hi0 lo1 hi1], of b: [lo2 hi2 lo3 hi3]
      lo[i] = __byte_perm(a, b, 0x6420);
      hi[i] = __byte_perm(a, b, 0x7531);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int32_t e0 = (int16_t)(w[2 * i + h] & 0xffff), e1 = (int16_t)(w[2 * i + h] >> 16);
        acc += (unsigned long long)(uint32_t)(e0 * e0) + (unsigned long long)(uint32_t)(e1 * e1);
      }
    }
    uint8_t *dst = limbs + row * ROWB + seg * 8;
    *reinterpret_cast<uint2 *>(dst) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2 *>(dst + 192) = make_uint2(lo[0], lo[1]);
    atomicAdd(&s_norm[r], acc);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int64_t rr = (int64_t)blockIdx.x * 8 + threadIdx.x;
    if (rr < n) {
      norms[rr] = (uint32_t)s_norm[threadIdx.x];   // mod 2^32, like the reference's Cardinal arithmetic
      // largest exact squared norm of the set (saturated): the top-k epilogue may compare with signed arithmetic only when
      // every distance provably stays below 2^31, i.e. every norm below 2^29
      if (norm_max) atomicMax(norm_max, (s_norm[threadIdx.x] >> 32) ? 0xFFFFFFFFu : (uint32_t)s_norm[threadIdx.x]);
    }
  }
}

constexpr int KMAX = 64;         // largest k
#ifndef TM_K1_NH
#define TM_K1_NH 2
#endif
#ifndef TM_STAG
#define TM_STAG 0
#endif

// ------------------------------------------------------------------ k = 1 kernel: 8 epilogue warps
// Same pipeline as above, but every TMEM lane quarter is drained by TWO warps (w and w + 4 may both address quarter
// w % 4), each taking half of the tile's 64 columns.  One warp per scheduler issues ~1 instruction every 3 cycles on
// this dependent integer code; two warps per scheduler hide each other's latencies.  A query row therefore has two
// threads, each with its own running arg-min; they are merged through shared memory at the end of the query block.
// NH column splits per tile: 4 NH epilogue warps + TMA warp + 2 MMA issuer warps
constexpr int k1_threads(int nh) { return (4 * nh + 3) * 32; }
// KS = 1: arg-min.  KS = 4: the four nearest, kept sorted in registers (k-means candidate search).
template <int KS, int NH>
__global__ void __launch_bounds__(k1_threads(NH), 1)
knn_i8_k1_kernel(const uint8_t *__restrict__ q_limbs, const __grid_constant__ CUtensorMap tmap_d,
                 const uint32_t *__restrict__ qnorm, const uint32_t *__restrict__ dnorm, int n_q, int n_dict,
                 int32_t *__restrict__ out_idx, uint32_t *__restrict__ out_dist, int tile_stride, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = STAGES_K1;
  uint8_t *sB = smem;
  constexpr int NEW = 4 * NH;   // epilogue warps; warp NEW = TMA producer, NEW + 1 / NEW + 2 = MMA issuers
  unsigned long long *s_merge = reinterpret_cast<unsigned long long *>(sB + NST * B_TILE);   // [NH - 1][KS][128]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_merge + BM * KS * (NH - 1));
  uint64_t *full = bars, *empty = bars + NST, *a_full = bars + 2 * NST, *a_empty = a_full + 1, *t_full = a_empty + 1,
           *t_empty = t_full + 2;
  uint64_t *stag = t_empty + 2;   // [4 quarters][NH splits][2 stages]: warp h of a quarter has its TMEM reads in flight
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(stag + 4 * NH * 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (n_dict + BN - 1) / BN;
  const int n_qblocks = (n_q + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 4);
    mbar_init(a_empty, 2);
    for (int g = 0; g < 2; ++g) { mbar_init(&t_full[g], 1); mbar_init(&t_empty[g], NEW); }
    for (int g = 0; g < 4 * NH * 2; ++g) mbar_init(&stag[g], 1);
    fence_barrier_init();
  }
  if (warp == NEW + 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == NEW && lane == 0) tma_prefetch_desc(&tmap_d);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0) __trap();   // the MMA issue code addresses TMEM with immediates

  if (warp == NEW) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x) {
        int jt = 0;
        for (int j = 0; j < n_tiles; ++j, ++it) {
          const uint32_t s = it % NST, r = it / NST;
          mbar_wait(&empty[s], (r & 1) ^ 1);
          if (dbg & 2) { mbar_arrive(&full[s]); jt += tile_stride; if (jt >= n_tiles) jt -= n_tiles; continue; }   // timing experiment: no dictionary traffic
          mbar_expect_tx(&full[s], B_TILE);
          for (int c = 0; c < 3; ++c) tma_load_2d(sB + s * B_TILE + c * CHUNK_B, &tmap_d, &full[s], c * 128, jt * BN);
          jt += tile_stride;
          if (jt >= n_tiles) jt -= n_tiles;
        }
      }
    }
  } else if (warp > NEW) {
    // ===================== two MMA issuer warps (even / odd tiles) =====================
    const uint32_t my_parity = (uint32_t)(warp - NEW - 1);
    constexpr uint32_t I_SS = make_idesc(kDFmtS32, kFmtS8, kFmtS8, BM, BN);
    constexpr uint32_t I_SU = make_idesc(kDFmtS32, kFmtS8, kFmtU8, BM, BN);
    constexpr uint32_t I_US = make_idesc(kDFmtS32, kFmtU8, kFmtS8, BM, BN);
    constexpr uint32_t I_UU = make_idesc(kDFmtS32, kFmtU8, kFmtU8, BM, BN);
    const uint32_t tA = tmem_base + A_COL;
    const uint64_t descB0 = umma_desc_sw128(smem_u32(sB));
    auto koffB = [](int ks) { return (uint64_t)(((ks >> 2) * CHUNK_B + (ks & 3) * 32) >> 4); };
    uint32_t it = 0, w = 0;
    for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
      mbar_wait(a_full, w & 1);
      tc_fence_after();
      for (int j = 0; j < n_tiles; ++j, ++it) {
        const uint32_t s = it % NST, r = it / NST;
        const uint32_t ts = it & 1;
        if (ts != my_parity) continue;
        mbar_wait(&full[s], r & 1);
        mbar_wait(&t_empty[ts], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dB = descB0 + (uint64_t)((s * B_TILE) >> 4);
        if (!(dbg & 4)) { if (my_parity == 0) mma_i8_tile_elect<0>((uint32_t)dB); else mma_i8_tile_elect<1>((uint32_t)dB); }
        tc_commit_elect(&t_full[ts]);
        tc_commit_elect(&empty[s]);
      }
      tc_commit_elect(a_empty);
    }
  } else {
    // ===================== epilogue: thread = (query row, column split h of NH) =====================
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int HN = BN / NH;   // columns per thread and tile
    uint32_t it = 0, w = 0;
    for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
      const int64_t qi = (int64_t)qb * BM + row;
      const bool valid = qi < n_q;
      const uint32_t nq = valid ? __ldg(qnorm + qi) : 0u;
      if (h == 0) {   // warps 0-3 also store the query rows into TMEM
        mbar_wait(a_empty, (w & 1) ^ 1);
        tc_fence_after();
        const uint4 *src = reinterpret_cast<const uint4 *>(q_limbs + (valid ? qi : 0) * ROWB);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          uint32_t r[16];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 t4 = valid ? __ldg(src + c * 4 + v) : make_uint4(0, 0, 0, 0);
            r[4 * v] = t4.x; r[4 * v + 1] = t4.y; r[4 * v + 2] = t4.z; r[4 * v + 3] = t4.w;
          }
          tmem_st16(t_lane + A_COL + c * 16, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      // the KS best (distance << 32 | index) keys of this thread's column half, ascending
      unsigned long long bk[KS];
#pragma unroll
      for (int r = 0; r < KS; ++r) bk[r] = ~0ull;
      uint32_t best_d = 0xFFFFFFFFu;   // distance of bk[KS - 1]: the admission threshold
      int jt = 0;
      for (int j = 0; j < n_tiles; ++j, ++it) {
        const uint32_t ts = it & 1;
        const int col0 = jt * BN + h * HN;
        jt += tile_stride;
        if (jt >= n_tiles) jt -= n_tiles;
        const int ncol = min(HN, n_dict - col0);   // may be <= 0 on the ragged last tile
        // dictionary norms of this tile half: independent of the MMA, so fetched before waiting for it
        uint32_t nd[HN];
#pragma unroll
        for (int v = 0; v < HN / 4; ++v) {
          const uint4 t4 = __ldg(reinterpret_cast<const uint4 *>(dnorm + col0) + v);
          nd[4 * v] = t4.x + nq; nd[4 * v + 1] = t4.y + nq; nd[4 * v + 2] = t4.z + nq; nd[4 * v + 3] = t4.w + nq;
        }
        mbar_wait(&t_full[ts], (it >> 1) & 1);
#if TM_STAG
        // The warps of a lane quarter take turns on the quarter's TMEM read port: warp h starts its reads only when warp
        // h - 1 has its own in flight (TM_STAG = 1) or complete (2).  Left alone the warps run in lock-step -- all read,
        // then all compute -- and the read port idles while the ALUs work and vice versa.
        if (h > 0) mbar_wait(&stag[((q * NH + h - 1) << 1) | ts], (it >> 1) & 1);
#endif
        tc_fence_after();
        const uint32_t t_acc = t_lane + ts * ACC_COLS + h * HN;
        // one burst: all three accumulators of this thread's 32 columns, then the stage goes straight back to the tensor
        // pipe (the TMEM hand-off chain, not the MMA rate, bounds a 2-stage ping-pong with K = 192)
        uint32_t pp[HN], xx[HN], lo[HN];
#pragma unroll
        for (int c = 0; c < HN / 16; ++c) {
          tmem_ld16(t_acc + c * 16, reinterpret_cast<uint32_t(&)[16]>(pp[c * 16]));
          tmem_ld16(t_acc + BN + c * 16, reinterpret_cast<uint32_t(&)[16]>(xx[c * 16]));
          tmem_ld16(t_acc + 2 * BN + c * 16, reinterpret_cast<uint32_t(&)[16]>(lo[c * 16]));
        }
#if TM_STAG == 1
        if (h < NH - 1 && lane == 0) mbar_arrive(&stag[((q * NH + h) << 1) | ts]);
#endif
        tmem_ld_wait();
#if TM_STAG == 2
        if (h < NH - 1 && lane == 0) mbar_arrive(&stag[((q * NH + h) << 1) | ts]);
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[ts]);
#pragma unroll
        for (int e = 0; e < HN; ++e) pp[e] = (pp[e] << 8) + xx[e];
#pragma unroll
        for (int ch = 0; ch < HN / 16; ++ch) {
          uint32_t dv[16];
          uint32_t m4[4];   // minimum of each group of four columns: the rare insertion path only walks groups that hold a candidate
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            m4[g] = 0xFFFFFFFFu;
#pragma unroll
            for (int e = 4 * g; e < 4 * g + 4; ++e) {
              uint32_t d = nd[ch * 16 + e];          // nq + nd
              d -= pp[ch * 16 + e] << 9;
              d -= lo[ch * 16 + e] << 1;
              dv[e] = d;
              m4[g] = min(m4[g], d);
            }
          }
          const int cbase = ch * 16;
          if (ncol < HN) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              m4[g] = 0xFFFFFFFFu;
#pragma unroll
              for (int e = 4 * g; e < 4 * g + 4; ++e) {
                if (cbase + e >= ncol) dv[e] = 0xFFFFFFFFu;
                m4[g] = min(m4[g], dv[e]);
              }
            }
          }
          const uint32_t m = min(min(m4[0], m4[1]), min(m4[2], m4[3]));
          // A warp takes this branch when ANY lane has a candidate: with the KS = 4 lists of the k-means search that is a quarter of
          // all 16-column units (32 lanes x 4 ln N insertions each), and walking all 16 columns every time made the k = 4 kernel
          // 1.6x slower per distance than k = 1.  Ties go to the lower dictionary index.
          if (m <= best_d) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (m4[g] > best_d) continue;
#pragma unroll
              for (int e = 4 * g; e < 4 * g + 4; ++e) {
                if (dv[e] == 0xFFFFFFFFu) continue;
                unsigned long long key = ((unsigned long long)dv[e] << 32) | (uint32_t)(col0 + cbase + e);
                if (key < bk[KS - 1]) {   // sorted insertion
#pragma unroll
                  for (int r = 0; r < KS; ++r) {
                    if (key < bk[r]) { const unsigned long long t = bk[r]; bk[r] = key; key = t; }
                  }
                  best_d = (uint32_t)(bk[KS - 1] >> 32);
                }
              }
            }
          }
        }
      }
      // merge the NH column splits of each row: the KS smallest of NH ascending lists
      if (h > 0) {
#pragma unroll
        for (int r = 0; r < KS; ++r) s_merge[((h - 1) * KS + r) * BM + row] = bk[r];
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(NEW * 32) : "memory");
      if (h == 0 && valid) {
#pragma unroll
        for (int o = 0; o < NH - 1; ++o) {
          unsigned long long ok[KS], mg[KS];
#pragma unroll
          for (int r = 0; r < KS; ++r) ok[r] = s_merge[(o * KS + r) * BM + row];
          int ia = 0, ib = 0;
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            // ia, ib <= r: static indexing through a small select keeps the lists in registers
            unsigned long long a = ~0ull, b = ~0ull;
#pragma unroll
            for (int u = 0; u < KS; ++u) { if (u == ia) a = bk[u]; if (u == ib) b = ok[u]; }
            const bool ta = a <= b;
            mg[r] = ta ? a : b;
            ia += ta; ib += !ta;
          }
#pragma unroll
          for (int r = 0; r < KS; ++r) bk[r] = mg[r];
        }
#pragma unroll
        for (int r = 0; r < KS; ++r) {
          out_idx[qi * KS + r] = (int32_t)(uint32_t)bk[r];        // 0xFFFFFFFF = -1 when fewer than KS rows exist
          out_dist[qi * KS + r] = (uint32_t)(bk[r] >> 32);
        }
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(NEW * 32) : "memory");   // s_merge is rewritten by the next query block
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NEW + 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ top-k kernel (2 <= k <= 64): 8 epilogue warps
// Same pipeline as the k = 1 kernel: every TMEM lane quarter is drained by two warps, each taking 32 of a tile's 64
// columns.  A thread = (query row, column half) runs its own streaming top-k over its 32 768-odd columns:
//   * admission is branch-free and PREDICATED: `d <= tau` guards one 8-byte store and one pointer bump, nothing else
//     (7-8 SASS instructions per distance, no shared-memory traffic, no serial select/move chain);
//   * the candidate list of a thread is a private 2 KB strip of a global workspace that stays in L2 (148 CTAs x 256
//     threads x 2 KB = 78 MB allocated, only admitted entries ever written);
//   * when a strip is nearly full the warp cuts it back with a LOOSE selection: bisection on the distance value until
//     between k and k + TK_SLACK entries survive (any threshold that keeps >= k entries is a valid tau).  That takes
//     ~6 warp reductions instead of the 32 of an exact radix select, and with TK_CAP = 256 a row is cut ~5 times per
//     scan instead of ~11;
//   * at the end of the query block the two strips of a row are merged by one exact selection under the
//     (distance, index) order -- the same order the oracle's brute force uses.
#ifndef TM_TK_CAP
#define TM_TK_CAP 256
#endif
constexpr int TK_CAP = TM_TK_CAP;  // candidate slots per (query row, column half)
constexpr int TK_SLACK = 16;     // a loose cut keeps between k and k + TK_SLACK candidates
constexpr int TK_CUT_FIRST = 4, TK_CUT_RATIO = 4;   // scheduled row-wide cuts after 4, 16, 64, 256, ... dictionary tiles
constexpr int TK_THREADS = 352;  // 8 epilogue warps + TMA warp + 2 MMA issuer warps
constexpr int STAGES_TK = 8;     // dictionary ring depth of the top-k kernel
// Ring of dictionary-norm slots (64 norms = 256 bytes per tile, copied by the TMA producer next to the tile).  A slot is read
// by the epilogue right after the TMEM read of its tile and rewritten STAGES_TK + 3 tiles later at the earliest (the
// producer runs at most STAGES_TK tiles ahead of MMA completion, the MMAs at most 2 tiles ahead of the TMEM reads).
constexpr int TK_NRING = STAGES_TK + 4;

__device__ __forceinline__ void lds_v4(uint32_t saddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(saddr));
}
__device__ __forceinline__ unsigned long long ldg_key(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg_key(unsigned long long *p, unsigned long long v) {
  asm volatile("st.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// if (d <= tau) { *(uint2 *)waddr = {idx, d}; waddr += 8; }   -- predicated, never a branch.  Only the low word of the
// pointer is bumped (a strip never straddles a 4 GB boundary), in place, so the pointer stays in one register pair.
__device__ __forceinline__ void tk_admit(unsigned long long &waddr, uint32_t idx, uint32_t d, uint32_t tau) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t"
      "setp.ls.u32 p, %2, %3;\n\t"
      "@p st.global.v2.u32 [%0], {%1, %2};\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "@p add.u32 lo, lo, 8;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}\n"
      : "+l"(waddr)
      : "r"(idx), "r"(d), "r"(tau)
      : "memory");
}
// d = nqnd - 2 (65536 HH + 256 X + LL)  (mod 2^32)
__device__ __forceinline__ uint32_t tk_dist(uint32_t nqnd, uint32_t hh, uint32_t xx, uint32_t ll) {
  const uint32_t a = hh * 256u + xx;
  const uint32_t b = a * 0xFFFFFE00u + nqnd;   // - 512 a
  return b - ll - ll;
}

// Whole warp; every lane holds NE (distance, index) entries, invalid ones as (0xFFFFFFFF, 0xFFFFFFFF); n_valid >= k.
// Finds T, TI such that the entries with  d < T  or  (d == T and index <= TI)  number between k and k + slack
// (exactly k when slack == 0) and contain the k smallest under the (distance, index) order.  Returns their count.
#ifdef TM_TK_TIMING
#define TKT_DECL long long tkt[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tkt_t = clock64_volatile();
#define TKT(i) { const long long t_ = clock64_volatile(); tkt[i] += t_ - tkt_t; tkt_t = t_; }
#define TKT_PRINT(role, n) if (blockIdx.x == 1 && lane == 0) printf("%s warp %d tiles %d: %lld %lld %lld %lld %lld %lld %lld %lld\n", role, warp, n, tkt[0] / (n), tkt[1] / (n), tkt[2] / (n), tkt[3] / (n), tkt[4] / (n), tkt[5] / (n), tkt[6] / (n), tkt[7] / (n));
__device__ long long g_cut_t[8];   // warp 0 of block 1: [loads, bisection, ties, compaction, cuts, probes]
#define CUT_ME (blockIdx.x == 1 && threadIdx.x == 0)
#define CUT_T(i) { const long long t_ = clock64_volatile(); if (CUT_ME) g_cut_t[i] += t_ - cut_t0; cut_t0 = t_; }
#define CUT_DECL long long cut_t0 = clock64_volatile();
#define CUT_ITER if (CUT_ME && slack) ++g_cut_t[5];
#define CUT_COUNT if (CUT_ME) ++g_cut_t[4];
#define CUT_PRINT if (CUT_ME) printf("cuts %lld probes %lld; cycles per cut: loads %lld bisect %lld ties %lld compact %lld\n", g_cut_t[4], g_cut_t[5], g_cut_t[0] / g_cut_t[4], g_cut_t[1] / g_cut_t[4], g_cut_t[2] / g_cut_t[4], g_cut_t[3] / g_cut_t[4]);
#else
#define TKT_DECL
#define TKT(i)
#define TKT_PRINT(role, n)
#define CUT_T(i)
#define CUT_DECL
#define CUT_ITER
#define CUT_COUNT
#define CUT_PRINT
#endif

// One 16-column unit of a tile: distances first (pure arithmetic, full ILP), then admission PAIR by PAIR behind a warp
// vote.  A memory instruction costs LSU issue bandwidth even when its predicate is false (32 predicated stores per
// tile and warp were ~40 % of the kernel), so the predicated store pair is only ISSUED when some lane of the warp
// admits one of the two columns -- late in the scan that is one pair in three.
__device__ __forceinline__ void tk_unit(unsigned long long &waddr, uint32_t tau, uint32_t nq, int col, int nvalid, const uint32_t (&nd)[16],
                                        uint32_t (&pp)[16], const uint32_t (&xx)[16], const uint32_t (&lo)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) pp[e] = tk_dist(nd[e] + nq, pp[e], xx[e], lo[e]);
  if (nvalid < 16) {   // ragged last dictionary tile: columns beyond the dictionary never compete
#pragma unroll
    for (int e = 0; e < 16; ++e) if (e >= nvalid) pp[e] = 0xFFFFFFFFu;
  }
  // The eight "does any lane admit pair e" questions of a unit are answered by ONE warp reduction (per-lane bit mask,
  // OR-reduced): eight vote -> branch chains in a row cost ~35 cycles of latency each.
  uint32_t mine = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) mine |= (min(pp[2 * e], pp[2 * e + 1]) <= tau) ? (1u << e) : 0u;
  const uint32_t any = __reduce_or_sync(0xffffffffu, mine);
  if (any != 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (any & (1u << e)) {
        tk_admit(waddr, (uint32_t)(col + 2 * e), pp[2 * e], tau);
        tk_admit(waddr, (uint32_t)(col + 2 * e + 1), pp[2 * e + 1], tau);
      }
    }
  }
}

// A whole (non-ragged) tile half of 32 columns: the 32 distances and the 16 pair masks form ONE basic block, so the
// multiply-adds of the distances (FMA pipe) and the min / compare / select work of the masks (ALU pipe) interleave in the
// instruction stream -- the two epilogue warps of a scheduler run in lock-step, and with the per-unit ordering both sat
// on the same pipe at the same time.  One warp reduction answers all 16 "does any lane admit pair e" questions.
__device__ __forceinline__ void tk_tile32(unsigned long long &waddr, uint32_t tau, uint32_t nq, int col, const uint32_t (&ndA)[16],
                                          uint32_t (&ppA)[16], const uint32_t (&xxA)[16], const uint32_t (&loA)[16],
                                          const uint32_t (&ndB)[16], uint32_t (&ppB)[16], const uint32_t (&xxB)[16],
                                          const uint32_t (&loB)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) ppA[e] = tk_dist(ndA[e] + nq, ppA[e], xxA[e], loA[e]);
#pragma unroll
  for (int e = 0; e < 16; ++e) ppB[e] = tk_dist(ndB[e] + nq, ppB[e], xxB[e], loB[e]);
  uint32_t mine = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) mine |= (min(ppA[2 * e], ppA[2 * e + 1]) <= tau) ? (1u << e) : 0u;
#pragma unroll
  for (int e = 0; e < 8; ++e) mine |= (min(ppB[2 * e], ppB[2 * e + 1]) <= tau) ? (0x100u << e) : 0u;
  const uint32_t any = __reduce_or_sync(0xffffffffu, mine);
  if (any != 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (any & (1u << e)) {
        tk_admit(waddr, (uint32_t)(col + 2 * e), ppA[2 * e], tau);
        tk_admit(waddr, (uint32_t)(col + 2 * e + 1), ppA[2 * e + 1], tau);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (any & (0x100u << e)) {
        tk_admit(waddr, (uint32_t)(col + 16 + 2 * e), ppB[2 * e], tau);
        tk_admit(waddr, (uint32_t)(col + 16 + 2 * e + 1), ppB[2 * e + 1], tau);
      }
    }
  }
}

// ---- no-wrap variant of tk_tile32 (every squared norm of the queries and of the dictionary is below 2^29, so every distance
// is below 2^31: true for any real feature set, |coefficient| <= 13 212).  The per-thread constant kk = |q|^2 - tau - 1 rides
// in the norm add, so  e = d - tau - 1  comes out of the same four instructions as d did, and "admit" is the SIGN of e: the
// pair mask is one OR and one funnel shift per column pair (2 instructions) instead of min / compare / select / add (3.3).
// Admitted entries store d = e + tau + 1.
__device__ __forceinline__ void tk_admit_nw(unsigned long long &waddr, uint32_t idx, uint32_t e, uint32_t tau1) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, d;\n\t"
      "setp.lt.s32 p, %2, 0;\n\t"
      "add.u32 d, %2, %3;\n\t"
      "@p st.global.v2.u32 [%0], {%1, d};\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "@p add.u32 lo, lo, 8;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}\n"
      : "+l"(waddr)
      : "r"(idx), "r"(e), "r"(tau1)
      : "memory");
}
// one column PAIR: both predicates first, the pointer bumped in place between the two predicated stores
__device__ __forceinline__ void tk_admit2_nw(unsigned long long &waddr, uint32_t idx, uint32_t e0, uint32_t e1, uint32_t tau1) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b32 lo, hi, d0, d1, i1;\n\t"
      "setp.lt.s32 p, %2, 0;\n\t"
      "setp.lt.s32 q, %3, 0;\n\t"
      "add.u32 d0, %2, %4;\n\t"
      "add.u32 d1, %3, %4;\n\t"
      "add.u32 i1, %1, 1;\n\t"
      "@p st.global.v2.u32 [%0], {%1, d0};\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "@p add.u32 lo, lo, 8;\n\t"
      "mov.b64 %0, {lo, hi};\n\t"
      "@q st.global.v2.u32 [%0], {i1, d1};\n\t"
      "@q add.u32 lo, lo, 8;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}\n"
      : "+l"(waddr)
      : "r"(idx), "r"(e0), "r"(e1), "r"(tau1)
      : "memory");
}
__device__ __forceinline__ void tk_tile32_nw(unsigned long long &waddr, uint32_t tau, uint32_t nq, int col, const uint32_t (&ndA)[16],
                                             uint32_t (&ppA)[16], const uint32_t (&xxA)[16], const uint32_t (&loA)[16],
                                             const uint32_t (&ndB)[16], uint32_t (&ppB)[16], const uint32_t (&xxB)[16],
                                             const uint32_t (&loB)[16]) {
  const uint32_t kk = nq - tau - 1u, tau1 = tau + 1u;
#pragma unroll
  for (int e = 0; e < 16; ++e) ppA[e] = tk_dist(ndA[e] + kk, ppA[e], xxA[e], loA[e]);
#pragma unroll
  for (int e = 0; e < 16; ++e) ppB[e] = tk_dist(ndB[e] + kk, ppB[e], xxB[e], loB[e]);
  uint32_t mine = 0;   // after the 16 shifts: bit 15 - p <-> pair p (p < 8: columns 2p, 2p + 1; p >= 8: columns 16 + 2(p - 8), ...)
#pragma unroll
  for (int e = 0; e < 8; ++e) mine = __funnelshift_l(ppA[2 * e] | ppA[2 * e + 1], mine, 1);
#pragma unroll
  for (int e = 0; e < 8; ++e) mine = __funnelshift_l(ppB[2 * e] | ppB[2 * e + 1], mine, 1);
  const uint32_t any = __reduce_or_sync(0xffffffffu, mine);
  if (any != 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (any & (0x8000u >> e)) tk_admit2_nw(waddr, (uint32_t)(col + 2 * e), ppA[2 * e], ppA[2 * e + 1], tau1);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (any & (0x80u >> e)) tk_admit2_nw(waddr, (uint32_t)(col + 16 + 2 * e), ppB[2 * e], ppB[2 * e + 1], tau1);
    }
  }
}

// half_out (optional): some value Th <= T_out with at least kh entries <= Th -- the first probe of the bisection whose count
// fell in [kh, k), for free; T_out itself when no probe did.
template <int NE>
__device__ __forceinline__ int tk_threshold(const uint32_t (&dd)[NE], const uint32_t (&ii)[NE], int n_valid, int k, int slack,
                                            uint32_t &T_out, uint32_t &TI_out, uint32_t *half_out = nullptr, int kh = 0) {
  uint32_t th = 0xFFFFFFFFu;
  uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    mn = min(mn, dd[i]);
    mx = max(mx, dd[i] == 0xFFFFFFFFu ? 0u : dd[i]);
  }
  uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
  int c_hi = n_valid;   // invariant: count(d <= hi) == c_hi >= k
  while (lo < hi && c_hi > k + slack) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    CUT_ITER
    int c = 0;
#pragma unroll
    for (int i = 0; i < NE; ++i) c += (dd[i] <= mid);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < k) { if (c >= kh && th == 0xFFFFFFFFu) th = mid; lo = mid + 1; }
    else { hi = mid; c_hi = c; }
  }
  T_out = hi;
  if (half_out) *half_out = min(th, hi);
  TI_out = 0xFFFFFFFFu;
  if (c_hi <= k + slack) return c_hi;
  // ties at the threshold distance: keep the (k - #{d < T}) lowest indices among them (rare)
  int cl = 0;
#pragma unroll
  for (int i = 0; i < NE; ++i) cl += (dd[i] < hi);
  cl = __reduce_add_sync(0xffffffffu, cl);
  const int need = k - cl;   // >= 1
  uint32_t TI = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t trial = TI | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < NE; ++i) c += (dd[i] == hi) && (ii[i] < trial);
    c = __reduce_add_sync(0xffffffffu, c);
    if (c < need) TI = trial;
  }
  TI_out = TI;   // need-th smallest index among the ties
  return k;
}

// Whole warp: cut the n candidates of the strip `buf` back to between k and k + TK_SLACK; returns the new count and
// the new admission threshold.
// other_half: the partner strip's published half-threshold (0xFFFFFFFF: none yet).  half_out: a value Th with at least
// ceil(k / 2) entries of THIS strip <= Th.  Two strips of a row cover disjoint columns, so max(Th, other_half) has at
// least k columns of the row below it and is a valid admission threshold for the whole row -- about the k-th distance over
// ALL columns seen so far instead of the k-th over this strip's half, which halves the steady-state admission rate.
__device__ __forceinline__ int tk_cut(unsigned long long *buf, int n, int k, int slack, int lane, uint32_t &tau_out, uint32_t other_half,
                                      uint32_t &half_out) {
  constexpr int NE = TK_CAP / 32;
  __syncwarp();
  CUT_DECL
  CUT_COUNT
  uint32_t dd[NE], ii[NE];
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const int p = i * 32 + lane;
    const unsigned long long key = p < n ? ldg_key(buf + p) : ~0ull;
    dd[i] = (uint32_t)(key >> 32);
    ii[i] = (uint32_t)key;
  }
  uint32_t T, TI;
  uint32_t Th;
#ifdef TM_TK_TIMING
  if (dd[0] == 1 && ii[NE - 1] == 12345) __trap();   // forces the loads to complete here
#endif
  CUT_T(0)
  const int kept = tk_threshold<NE>(dd, ii, n, k, slack, T, TI, &Th, (k + 1) >> 1);
  CUT_T(1)
  half_out = Th;
  const uint32_t T_row = max(Th, other_half);
  if (T_row < T) { T = T_row; TI = 0xFFFFFFFFu; }   // the row-wide threshold is tighter: keep every entry <= it
  __syncwarp();   // every lane has its entries in registers before the strip is rewritten
  const uint32_t lt_mask = (1u << lane) - 1u;
  int outp = 0;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool keep = (dd[i] < T) || (dd[i] == T && ii[i] <= TI && dd[i] != 0xFFFFFFFFu);
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    if (keep) stg_key(buf + outp + __popc(km & lt_mask), ((unsigned long long)dd[i] << 32) | ii[i]);
    outp += __popc(km);
  }
  __syncwarp();
  CUT_T(3)
  (void)kept;
  tau_out = T;
  return outp;
}

// Whole warp: ROW-WIDE cut of the two strips of one query row (b0: column half 0, b1: column half 1; n0 / n1 entries).
// Finds T leaving between k and k + slack entries of the UNION at or below it (a valid admission threshold for every later
// column of the row, and ~sqrt(2) tighter in rank than what either strip could certify alone), compacts each strip in
// place and returns the new counts.  With fewer than k entries in the union nothing can be dropped.
// Scheduled cuts call this for every row of the CTA at the same dictionary tile, so no warp waits for another warp's cut,
// and two rows are in flight per call site (the loads of the second row overlap the selection of the first).
__device__ __forceinline__ void tk_row_load(const unsigned long long *b0, const unsigned long long *b1, int n0, int n1, int lane,
                                            uint32_t (&dd)[2 * (TK_CAP / 32)], uint32_t (&ii)[2 * (TK_CAP / 32)]) {
  constexpr int NE = TK_CAP / 32;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const int p = i * 32 + lane;
    const unsigned long long k0 = p < n0 ? ldg_key(b0 + p) : ~0ull, k1 = p < n1 ? ldg_key(b1 + p) : ~0ull;
    dd[i] = (uint32_t)(k0 >> 32); ii[i] = (uint32_t)k0;
    dd[NE + i] = (uint32_t)(k1 >> 32); ii[NE + i] = (uint32_t)k1;
  }
}
__device__ __forceinline__ uint32_t tk_row_cut(unsigned long long *b0, unsigned long long *b1, int n0, int n1, int k, int slack, int lane,
                                               const uint32_t (&dd)[2 * (TK_CAP / 32)], const uint32_t (&ii)[2 * (TK_CAP / 32)], int &c0,
                                               int &c1) {
  constexpr int NE = TK_CAP / 32;
  c0 = n0; c1 = n1;
  if (n0 + n1 <= k) return 0xFFFFFFFEu;
  uint32_t T, TI;
  tk_threshold<2 * NE>(dd, ii, n0 + n1, k, slack, T, TI);
  const uint32_t lt_mask = (1u << lane) - 1u;
  int o0 = 0, o1 = 0;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool keep = (dd[i] < T) || (dd[i] == T && ii[i] <= TI && dd[i] != 0xFFFFFFFFu);
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    if (keep) stg_key(b0 + o0 + __popc(km & lt_mask), ((unsigned long long)dd[i] << 32) | ii[i]);
    o0 += __popc(km);
  }
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool keep = (dd[NE + i] < T) || (dd[NE + i] == T && ii[NE + i] <= TI && dd[NE + i] != 0xFFFFFFFFu);
    const uint32_t km = __ballot_sync(0xffffffffu, keep);
    if (keep) stg_key(b1 + o1 + __popc(km & lt_mask), ((unsigned long long)dd[NE + i] << 32) | ii[NE + i]);
    o1 += __popc(km);
  }
  c0 = o0; c1 = o1;
  return T;
}

__global__ void __launch_bounds__(TK_THREADS, 1)
knn_i8_topk_kernel(const uint8_t *__restrict__ q_limbs, const __grid_constant__ CUtensorMap tmap_d,
                   const uint32_t *__restrict__ qnorm, const uint32_t *__restrict__ dnorm, int n_q, int n_dict, int k,
                   int32_t *__restrict__ out_idx, uint32_t *__restrict__ out_dist, int tile_stride,
                   unsigned long long *ws /* [gridDim.x][256][TK_CAP], 2 KB aligned */, int slack, int dbg, int cut_first,
                   int cut_ratio /* scheduled row-wide cuts after cut_first, cut_first * cut_ratio, ... tiles; 0 = none */,
                   const uint32_t *__restrict__ q_nmax, const uint32_t *__restrict__ d_nmax /* largest squared norms, or null */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = STAGES_TK;
  uint8_t *sB = smem;
  int32_t *s_cnt = reinterpret_cast<int32_t *>(sB + NST * B_TILE);    // [256] candidates per thread (row-wide cuts, end of a query block)
  uint32_t *s_thalf = reinterpret_cast<uint32_t *>(s_cnt + 256);      // [2][128] half-thresholds published by the strips of a row
  uint32_t *s_tau = s_thalf + 256;                                    // [128] row-wide threshold left by a scheduled cut
  uint32_t *s_nd = s_tau + 128;                                       // [TK_NRING][64] dictionary norms of the tiles in flight
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_nd + TK_NRING * BN);
  uint64_t *full = bars, *empty = bars + NST, *a_full = bars + 2 * NST, *a_empty = a_full + 1, *t_full = a_empty + 1,
           *t_empty = t_full + 2;
  uint64_t *stag = t_empty + 2;   // [4 quarters][2 stages]: the first warp of a quarter has its TMEM reads in flight
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(stag + 8);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int n_tiles = (n_dict + BN - 1) / BN;
  const int n_qblocks = (n_q + BM - 1) / BM;

  if (threadIdx.x < 256) s_thalf[threadIdx.x] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 4);
    mbar_init(a_empty, 2);
    for (int g = 0; g < 2; ++g) { mbar_init(&t_full[g], 1); mbar_init(&t_empty[g], 8); }
    for (int g = 0; g < 8; ++g) mbar_init(&stag[g], 1);
    fence_barrier_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tmap_d);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0) __trap();   // the MMA issue code addresses TMEM with immediates

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x) {
        int jt = 0;
        for (int j = 0; j < n_tiles; ++j, ++it) {
          const uint32_t s = it % NST, r = it / NST;
          mbar_wait(&empty[s], (r & 1) ^ 1);
          if (dbg & 2) { mbar_arrive(&full[s]); jt += tile_stride; if (jt >= n_tiles) jt -= n_tiles; continue; }   // timing experiment: no dictionary traffic
          mbar_expect_tx(&full[s], B_TILE + BN * 4);
          for (int c = 0; c < 3; ++c) tma_load_2d(sB + s * B_TILE + c * CHUNK_B, &tmap_d, &full[s], c * 128, jt * BN);
          bulk_load_1d(s_nd + (it % TK_NRING) * BN, dnorm + (size_t)jt * BN, BN * 4, &full[s]);   // norms are padded to whole tiles
          jt += tile_stride;
          if (jt >= n_tiles) jt -= n_tiles;
        }
      }
    }
  } else if (warp >= 9) {
    // ===================== two MMA issuer warps (even / odd tiles) =====================
    const uint32_t my_parity = (uint32_t)(warp - 9);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(sB));
    uint32_t it = 0, w = 0;
    TKT_DECL
    for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
      mbar_wait(a_full, w & 1);
      tc_fence_after();
      TKT(0)
      for (int j = 0; j < n_tiles; ++j, ++it) {
        const uint32_t s = it % NST, r = it / NST;
        const uint32_t ts = it & 1;
        if (ts != my_parity) continue;
        mbar_wait(&full[s], r & 1);
        TKT(1)
        mbar_wait(&t_empty[ts], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        TKT(2)
        const uint64_t dB = descB0 + (uint64_t)((s * B_TILE) >> 4);
        if (!(dbg & 4)) { if (my_parity == 0) mma_i8_tile_elect<0>((uint32_t)dB); else mma_i8_tile_elect<1>((uint32_t)dB); }
        tc_commit_elect(&t_full[ts]);
        tc_commit_elect(&empty[s]);
        TKT(3)
      }
      tc_commit_elect(a_empty);
    }
    TKT_PRINT("mma [a_full, full, t_empty, issue]", (int)(it / 2))
  } else {
    // ===================== epilogue: thread = (query row, column half) =====================
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int HN = BN / 2;   // 32 columns per thread and tile
    unsigned long long *cta_ws = ws + (size_t)blockIdx.x * 256 * TK_CAP;
    unsigned long long *wbuf = cta_ws + (size_t)(warp * 32) * TK_CAP;   // this warp's 32 strips
    unsigned long long *mybuf = wbuf + (size_t)lane * TK_CAP;
    const uint32_t base_lo = (uint32_t)(uintptr_t)mybuf;
    // signed comparisons are exact when no distance can reach 2^31: every squared norm below 2^29 ((|q| + |t|)^2 < 2^31)
    const bool nowrap = !(dbg & 64) && q_nmax && d_nmax && __ldg(q_nmax) < (1u << 29) && __ldg(d_nmax) < (1u << 29);
    uint32_t it = 0, w = 0;
    int nslot = 0;                                   // it % TK_NRING, kept as a counter
    const uint32_t s_nd_u32 = smem_u32(s_nd);
    TKT_DECL
    for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
      const int64_t qi = (int64_t)qb * BM + row;
      const bool valid = qi < n_q;
      const uint32_t nq = valid ? __ldg(qnorm + qi) : 0u;
      if (h == 0) {   // warps 0-3 also store the query rows into TMEM
        mbar_wait(a_empty, (w & 1) ^ 1);
        tc_fence_after();
        const uint4 *src = reinterpret_cast<const uint4 *>(q_limbs + (valid ? qi : 0) * ROWB);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          uint32_t r[16];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 t4 = valid ? __ldg(src + c * 4 + v) : make_uint4(0, 0, 0, 0);
            r[4 * v] = t4.x; r[4 * v + 1] = t4.y; r[4 * v + 2] = t4.z; r[4 * v + 3] = t4.w;
          }
          tmem_st16(t_lane + A_COL + c * 16, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
      // distances of 0xFFFFFFFF (masked columns) are never admitted; without wrap every distance is below 2^31
      uint32_t tau = (dbg & 1) ? 0u : (nowrap ? 0x7FFFFFFEu : 0xFFFFFFFEu);
      unsigned long long waddr = (unsigned long long)(uintptr_t)mybuf;   // next free slot of this thread's strip
      uint32_t my_half = 0xFFFFFFFFu;                                     // this strip's published half-threshold
      int jt = 0;
      int next_cut = cut_first > 0 ? cut_first : 0x7fffffff;
      TKT(0)
      for (int j = 0; j < n_tiles; ++j, ++it) {
        const uint32_t ts = it & 1;
        const int col0 = jt * BN + h * HN;
        jt += tile_stride;
        if (jt >= n_tiles) jt -= n_tiles;
        TKT(1)
        mbar_wait(&t_full[ts], (it >> 1) & 1);
#if TM_STAG
        if (h > 0) mbar_wait(&stag[(q << 1) | ts], (it >> 1) & 1);   // see knn_i8_k1_kernel: the quarter's two warps take turns on the read port
#endif
        tc_fence_after();
        TKT(2)
        const uint32_t t_acc = t_lane + ts * ACC_COLS + h * HN;
        uint32_t ppA[16], xxA[16], loA[16], ppB[16], xxB[16], loB[16];
        tmem_ld16(t_acc, ppA);
        tmem_ld16(t_acc + BN, xxA);
        tmem_ld16(t_acc + 2 * BN, loA);
        tmem_ld16(t_acc + 16, ppB);
        tmem_ld16(t_acc + BN + 16, xxB);
        tmem_ld16(t_acc + 2 * BN + 16, loB);
#if TM_STAG == 1
        if (h == 0 && lane == 0) mbar_arrive(&stag[(q << 1) | ts]);
#endif
        // dictionary norms of this tile half from the ring the TMA producer fills (broadcast reads; they overlap the TMEM reads)
        uint32_t ndA[16], ndB[16];
        {
          const uint32_t nsrc = s_nd_u32 + (uint32_t)(nslot * BN + h * HN) * 4u;   // explicit ld.shared: a generic load costs an address-space check
          if (++nslot == TK_NRING) nslot = 0;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            lds_v4(nsrc + 16 * v, ndA[4 * v], ndA[4 * v + 1], ndA[4 * v + 2], ndA[4 * v + 3]);
            lds_v4(nsrc + 64 + 16 * v, ndB[4 * v], ndB[4 * v + 1], ndB[4 * v + 2], ndB[4 * v + 3]);
          }
        }
        tmem_ld_wait();
#if TM_STAG == 2
        if (h == 0 && lane == 0) mbar_arrive(&stag[(q << 1) | ts]);
#endif
        tc_fence_before();          // the whole tile is in registers: hand the stage back to the tensor pipe
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[ts]);
        TKT(3)
        if (!(dbg & 8)) {
          if (col0 + HN <= n_dict) {
            if (nowrap) tk_tile32_nw(waddr, tau, nq, col0, ndA, ppA, xxA, loA, ndB, ppB, xxB, loB);
            else tk_tile32(waddr, tau, nq, col0, ndA, ppA, xxA, loA, ndB, ppB, xxB, loB);
          } else {   // ragged last dictionary tile
            tk_unit(waddr, tau, nq, col0, n_dict - col0, ndA, ppA, xxA, loA);
            tk_unit(waddr, tau, nq, col0 + 16, n_dict - col0 - 16, ndB, ppB, xxB, loB);
          }
        }
        TKT(4)
        TKT(5)
        // the next check is a tile (HN admissions at most) away
        const uint32_t wlo = (uint32_t)waddr;
        uint32_t fullm = __ballot_sync(0xffffffffu, wlo - base_lo > (uint32_t)((TK_CAP - HN) * 8));
        while (fullm) {
          const int L = __ffs(fullm) - 1;
          fullm &= fullm - 1;
          const int nL = (int)((__shfl_sync(0xffffffffu, wlo, L) - __shfl_sync(0xffffffffu, base_lo, L)) >> 3);
          uint32_t T, Th;
          const uint32_t other = (dbg & 16) ? 0xFFFFFFFFu : *(volatile uint32_t *)(s_thalf + (h ^ 1) * 128 + q * 32 + L);
          const int kept = tk_cut(wbuf + (size_t)L * TK_CAP, nL, k, slack, lane, T, other, Th);
          if (lane == L) {
            waddr = (unsigned long long)(uintptr_t)(mybuf + kept); tau = T; my_half = Th;
            *(volatile uint32_t *)(s_thalf + h * 128 + row) = Th;
          }
        }
        // ---- scheduled ROW-WIDE cut: every warp of the CTA cuts all its rows after the same dictionary tile
        if (j + 1 == next_cut) {
          next_cut = (cut_ratio > 1 && (long long)next_cut * cut_ratio < n_tiles) ? next_cut * cut_ratio : 0x7fffffff;
          s_cnt[warp * 32 + lane] = (int)(((uint32_t)waddr - base_lo) >> 3);
          asm volatile("bar.sync %0, 64;\n" ::"r"(2 + q) : "memory");   // the two warps of this lane quarter: strips and counts visible
          // warp (q, h) cuts rows [16 h, 16 h + 16) of the quarter, two at a time
          for (int rr = 16 * h; rr < 16 * h + 16; rr += 2) {
            const int r0 = q * 32 + rr, r1 = r0 + 1;
            unsigned long long *a0 = cta_ws + (size_t)r0 * TK_CAP, *a1 = cta_ws + (size_t)(128 + r0) * TK_CAP;
            unsigned long long *c0p = cta_ws + (size_t)r1 * TK_CAP, *c1p = cta_ws + (size_t)(128 + r1) * TK_CAP;
            const int na0 = s_cnt[r0], na1 = s_cnt[128 + r0], nc0 = s_cnt[r1], nc1 = s_cnt[128 + r1];
            uint32_t ddA[2 * (TK_CAP / 32)], iiA[2 * (TK_CAP / 32)], ddC[2 * (TK_CAP / 32)], iiC[2 * (TK_CAP / 32)];
            tk_row_load(a0, a1, na0, na1, lane, ddA, iiA);
            tk_row_load(c0p, c1p, nc0, nc1, lane, ddC, iiC);
            int k0, k1;
            const uint32_t TA = tk_row_cut(a0, a1, na0, na1, k, slack, lane, ddA, iiA, k0, k1);
            if (lane == 0) { s_cnt[r0] = k0; s_cnt[128 + r0] = k1; s_tau[r0] = TA; }
            const uint32_t TC = tk_row_cut(c0p, c1p, nc0, nc1, k, slack, lane, ddC, iiC, k0, k1);
            if (lane == 0) { s_cnt[r1] = k0; s_cnt[128 + r1] = k1; s_tau[r1] = TC; }
          }
          asm volatile("bar.sync %0, 64;\n" ::"r"(2 + q) : "memory");
          waddr = (unsigned long long)(uintptr_t)(mybuf + s_cnt[warp * 32 + lane]);
          tau = min(tau, s_tau[row]);
        }
        TKT(6)
      }
      // ---- results of this query block: merge the two column halves of every row
      s_cnt[warp * 32 + lane] = (int)(((uint32_t)waddr - base_lo) >> 3);
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      s_thalf[h * 128 + row] = 0xFFFFFFFFu;   // every scan of this block is over; the barrier below publishes the reset
      {
        constexpr int NE = 2 * TK_CAP / 32;
        const uint32_t lt_mask = (1u << lane) - 1u;
        for (int L = 16 * h; L < 16 * h + 16; ++L) {   // warp (q, h) merges rows [16 h, 16 h + 16) of quarter q
          const int64_t qL = (int64_t)qb * BM + q * 32 + L;
          if (qL >= n_q) break;
          const int n0 = s_cnt[q * 32 + L], n1 = s_cnt[128 + q * 32 + L];
          const unsigned long long *b0 = cta_ws + (size_t)(q * 32 + L) * TK_CAP, *b1 = cta_ws + (size_t)(128 + q * 32 + L) * TK_CAP;
          uint32_t dd[NE], ii[NE];
#pragma unroll
          for (int i = 0; i < NE / 2; ++i) {
            const int p = i * 32 + lane;
            const unsigned long long k0 = p < n0 ? ldg_key(b0 + p) : ~0ull, k1 = p < n1 ? ldg_key(b1 + p) : ~0ull;
            dd[i] = (uint32_t)(k0 >> 32); ii[i] = (uint32_t)k0;
            dd[NE / 2 + i] = (uint32_t)(k1 >> 32); ii[NE / 2 + i] = (uint32_t)k1;
          }
          uint32_t T = 0xFFFFFFFEu, TI = 0xFFFFFFFFu;
          if (n0 + n1 > k) tk_threshold<NE>(dd, ii, n0 + n1, k, 0, T, TI);
          int outp = 0;
#pragma unroll
          for (int i = 0; i < NE; ++i) {
            const bool keep = (dd[i] < T) || (dd[i] == T && ii[i] <= TI && dd[i] != 0xFFFFFFFFu);
            const uint32_t km = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const int pos = outp + __popc(km & lt_mask);
              out_idx[qL * k + pos] = (int32_t)ii[i];
              out_dist[qL * k + pos] = dd[i];
            }
            outp += __popc(km);
          }
          for (int pos = outp + lane; pos < k; pos += 32) {   // fewer than k dictionary rows: empty slots are (-1, 0xFFFFFFFF)
            out_idx[qL * k + pos] = -1;
            out_dist[qL * k + pos] = 0xFFFFFFFFu;
          }
        }
      }
      asm volatile("bar.sync 1, 256;\n" ::: "memory");   // strips and s_cnt are reused by the next query block
      TKT(7)
    }
    TKT_PRINT("epi [qblock setup, nd loads, t_full wait, ld+release, unit A, unit B, cuts, merge]", (int)it)
    CUT_PRINT
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ row sort: (distance, index) ascending, k <= 64
__global__ void __launch_bounds__(256) knn_sort_rows_kernel(int32_t *__restrict__ idx, uint32_t *__restrict__ dist, int64_t n_q, int k) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_q) return;
  uint64_t key[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = i * 32 + lane;
    key[i] = p < k ? ((uint64_t)dist[row * k + p] << 32) | (uint32_t)idx[row * k + p] : ~0ull;
  }
  for (int size = 2; size <= 64; size <<= 1) {
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      if (stride == 32) {
        // partner is the other register of the same lane; size == 64 -> ascending everywhere
        if (key[0] > key[1]) { const uint64_t t = key[0]; key[0] = key[1]; key[1] = t; }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int e = i * 32 + lane;
          const uint64_t other = __shfl_xor_sync(0xffffffffu, key[i], stride);
          const bool asc = (e & size) == 0;
          const bool lower = (e & stride) == 0;
          const bool take_min = (asc == lower);
          key[i] = take_min ? (key[i] < other ? key[i] : other) : (key[i] > other ? key[i] : other);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = i * 32 + lane;
    if (p < k) { dist[row * k + p] = (uint32_t)(key[i] >> 32); idx[row * k + p] = (int32_t)(uint32_t)key[i]; }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// uint8 tensor [rows][row_bytes] row-major, box = 128 bytes x box_rows, 128-byte swizzle
int make_tmap_rows_u8(CUtensorMap *map, const void *base, uint64_t rows, uint32_t row_bytes, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return TM_ERR_DRIVER;
  cuuint64_t dims[2] = {row_bytes, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TM_OK : TM_ERR_DRIVER;
}

int knn_rows_per_cta() { return BM; }

int launch_limb_split(const int16_t *in, int64_t n, uint8_t *limbs, uint32_t *norms, cudaStream_t st, uint32_t *norm_max) {
  if (n <= 0) return TM_OK;
  limb_split_kernel<<<(unsigned)((n + 7) / 8), 192, 0, st>>>(in, n, limbs, norms, norm_max);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_knn_i8(const uint8_t *q_limbs, const uint32_t *q_norm, int n_q, const uint8_t *d_limbs, const uint32_t *d_norm,
                  int n_dict, int k, int32_t *out_idx, uint32_t *out_dist, void *ws, int num_ctas, int sort_rows,
                  cudaStream_t st, const uint32_t *q_norm_max, const uint32_t *d_norm_max) {
  (void)ws;
  if (n_q <= 0) return TM_OK;
  if (k < 1 || k > KMAX || n_dict <= 0) return TM_ERR_ARG;
  CUtensorMap td;
  int rc = make_tmap_rows_u8(&td, d_limbs, (uint64_t)n_dict, ROWB, BN);
  if (rc != TM_OK) return rc;
  constexpr int K1_NH = TM_K1_NH;   // column splits per tile in the k = 1 / k = 4 kernels (2: 8 epilogue warps, 4: 16)
  constexpr int SMEM_K1 = STAGES_K1 * B_TILE + BM * 8 * 4 * (K1_NH - 1) + 768 + 1024;
  constexpr int SMEM_TK = STAGES_TK * B_TILE + 256 * 4 + 256 * 4 + 128 * 4 + TK_NRING * BN * 4 + 768 + 1024;
  static bool attr_set[TM_MAX_DEVICES] = {};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(knn_i8_k1_kernel<1, K1_NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_K1) != cudaSuccess) return TM_ERR_CUDA;
    if (cudaFuncSetAttribute(knn_i8_k1_kernel<4, K1_NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_K1) != cudaSuccess) return TM_ERR_CUDA;
    if (cudaFuncSetAttribute(knn_i8_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TK) != cudaSuccess) return TM_ERR_CUDA;
  }
  {
    ProfScope prof(k == 1 ? "knn_k1" : (k == 4 ? "knn_k4" : "knn_topk"), st);
    // Dictionary tiles are visited in the order j -> (j * stride) mod n_tiles with stride ~ 0.618 n_tiles, coprime to
    // n_tiles: a streaming top-k admits ~k ln(N/k) candidates per row when the order is uncorrelated with the distance,
    // but nearly all N when distances fall along the scan (dictionaries built frame by frame do exactly that).
    const int n_tiles_h = (n_dict + BN - 1) / BN;
    int tile_stride = (int)(0.6180339887 * n_tiles_h);
    if (tile_stride < 1) tile_stride = 1;
    auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
    while (gcd(tile_stride, n_tiles_h) != 1) ++tile_stride;
    const int n_qblocks = (n_q + BM - 1) / BM;
    const int grid = n_qblocks < num_ctas ? n_qblocks : num_ctas;
    static int kdbg = -1;   // TM_TK_DBG bits: 1 admit nothing (top-k), 2 no dictionary loads, 4 no MMAs, 8 no epilogue arithmetic, 16 no threshold sharing, 32 staged admission
    if (kdbg < 0) kdbg = getenv("TM_TK_DBG") ? atoi(getenv("TM_TK_DBG")) : 0;
    if (k == 1) knn_i8_k1_kernel<1, K1_NH><<<grid, k1_threads(K1_NH), SMEM_K1, st>>>(q_limbs, td, q_norm, d_norm, n_q, n_dict, out_idx, out_dist, tile_stride, kdbg);
    else if (k == 4) knn_i8_k1_kernel<4, K1_NH><<<grid, k1_threads(K1_NH), SMEM_K1, st>>>(q_limbs, td, q_norm, d_norm, n_q, n_dict, out_idx, out_dist, tile_stride, kdbg);
    else {
      // candidate strips: stream-ordered scratch (the pool keeps it cached between calls); 2 KB alignment keeps every
      // strip inside one 4 GB window, so the kernel bumps only the low word of its write pointer
      static_assert((TK_CAP * 8) % 2048 == 0, "a strip must not straddle a 4 GB boundary");
      void *raw = nullptr;
      const size_t strips = (size_t)grid * 256 * TK_CAP * 8;
      if (cudaMallocAsync(&raw, strips + 2048, st) != cudaSuccess) return TM_ERR_NOMEM;
      unsigned long long *strip_ws = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(raw) + 2047) & ~uintptr_t(2047));
      static int slack = -1, dbg = 0, cut_first = TK_CUT_FIRST, cut_ratio = TK_CUT_RATIO;   // TM_TK_SLACK / TM_TK_DBG / TM_TK_CUTS="first,ratio": tuning and timing experiments
      if (slack < 0) {
        slack = getenv("TM_TK_SLACK") ? atoi(getenv("TM_TK_SLACK")) : TK_SLACK;
        dbg = getenv("TM_TK_DBG") ? atoi(getenv("TM_TK_DBG")) : 0;
        if (const char *e = getenv("TM_TK_CUTS")) { if (sscanf(e, "%d,%d", &cut_first, &cut_ratio) != 2) { cut_first = TK_CUT_FIRST; cut_ratio = TK_CUT_RATIO; } }
      }
      knn_i8_topk_kernel<<<grid, TK_THREADS, SMEM_TK, st>>>(q_limbs, td, q_norm, d_norm, n_q, n_dict, k, out_idx, out_dist, tile_stride, strip_ws, slack, dbg,
                                                            cut_first, cut_ratio, q_norm_max, d_norm_max);
      cudaFreeAsync(raw, st);
    }
  }
  note_launch();
  if (cudaGetLastError() != cudaSuccess) return TM_ERR_CUDA;
  if (k > 1 && k != 4 && sort_rows) {   // the k = 4 kernel emits sorted rows
    note_launch();
    knn_sort_rows_kernel<<<(unsigned)((n_q + 7) / 8), 256, 0, st>>>(out_idx, out_dist, n_q, k);
    if (cudaGetLastError() != cudaSuccess) return TM_ERR_CUDA;
  }
  return TM_OK;
}

}  // namespace tmg
