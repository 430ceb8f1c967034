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
// CTA (persistent, one per SM): 256 queries (two 128-row M tiles, resident in smem) x the whole dictionary
// streamed in 64-row N tiles through a 4-stage TMA ring.  M tile g accumulates into TMEM stage g (3 x 64 columns:
// HH, X, LL), so the tensor pipe works on tile g^1 while the four epilogue warps of tile g fold the accumulators
// into distances and keep, per query row (one thread = one TMEM lane = one query), either the running arg-min
// (k = 1) or a thresholded candidate list that is cut back to the k best by a warp-cooperative radix select.
#include "tc_common.cuh"
#include "tm_kernels.h"

namespace tmg {

constexpr int BM = 128;          // rows per M tile (UMMA M)
constexpr int BN = 64;           // rows per N tile (UMMA N)
constexpr int STAGES = 4;        // dictionary ring depth
constexpr int ROWB = 384;        // bytes per limb row
constexpr int CHUNK_A = BM * 128;  // one 128-byte-wide swizzle chunk of an A tile
constexpr int CHUNK_B = BN * 128;
constexpr int A_TILE = 3 * CHUNK_A;  // 49152
constexpr int B_TILE = 3 * CHUNK_B;  // 24576
constexpr int ACC_COLS = 3 * BN; // TMEM columns per stage

// ------------------------------------------------------------------ limb split + norms
// in: [n][192] int16 -> limbs [n][384] (hi bytes then lo bytes), norms[n] = sum v^2 mod 2^32
__global__ void __launch_bounds__(192) limb_split_kernel(const int16_t *__restrict__ in, int64_t n, uint8_t *__restrict__ limbs,
                                                        uint32_t *__restrict__ norms) {
  __shared__ uint32_t s_norm[8];
  const int r = threadIdx.x / 24, seg = threadIdx.x % 24;
  const int64_t row = (int64_t)blockIdx.x * 8 + r;
  if (threadIdx.x < 8) s_norm[threadIdx.x] = 0;
  __syncthreads();
  if (row < n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + row * 192) + seg);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[2], lo[2], acc = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t a = w[2 * i], b = w[2 * i + 1];
      // bytes of a: [lo0 hi0 lo1 hi1], of b: [lo2 hi2 lo3 hi3]
      lo[i] = __byte_perm(a, b, 0x6420);
      hi[i] = __byte_perm(a, b, 0x7531);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int32_t e0 = (int16_t)(w[2 * i + h] & 0xffff), e1 = (int16_t)(w[2 * i + h] >> 16);
        acc += (uint32_t)(e0 * e0) + (uint32_t)(e1 * e1);
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
    if (rr < n) norms[rr] = s_norm[threadIdx.x];
  }
}

// ------------------------------------------------------------------ top-k state in shared memory
// Per query row (= per epilogue thread): a max-heap of the k best (distance << 32 | index) keys and a small queue of
// admitted-but-not-yet-inserted candidates.  Arrays are [slot][row] so a warp touches 32 consecutive words whatever
// slot each lane is at.  Insertions are deferred and done by all lanes together (the sift-down loop is divergent:
// batching makes every trip through it serve many rows at once).
constexpr int KMAX = 64;         // heap slots per row
constexpr int QCAP = 16;         // queue slots per row
constexpr int TK_ROWS = 128;     // rows per CTA on the top-k path

__device__ __forceinline__ void heap_replace_root(unsigned long long *heap, int row, int k, unsigned long long key) {
  // heap[slot * TK_ROWS + row]; root holds the largest key; key < root guaranteed by the caller
  int i = 0;
  for (;;) {
    const int l = 2 * i + 1;
    if (l >= k) break;
    int c = l;
    unsigned long long ck = heap[l * TK_ROWS + row];
    if (l + 1 < k) {
      const unsigned long long rk = heap[(l + 1) * TK_ROWS + row];
      if (rk > ck) { ck = rk; c = l + 1; }
    }
    if (ck <= key) break;
    heap[i * TK_ROWS + row] = ck;
    i = c;
  }
  heap[i * TK_ROWS + row] = key;
}

// ------------------------------------------------------------------ main kernel
// MT = M tiles (128 query rows each) per CTA.  Work is a flat sequence of units u = (dictionary tile j, M tile g),
// u = j*MT + g; unit u accumulates into TMEM stage u & 1.  MT = 2: epilogue warps 0-3 own M tile 0 / stage 0, warps
// 4-7 own M tile 1 / stage 1 (k = 1 path).  MT = 1: the four epilogue warps alternate between the two stages (top-k
// path, which needs the shared memory for the heaps).
template <int MT, bool TOPK>
__global__ void __launch_bounds__(64 + 128 * MT, 1)
knn_i8_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
              const uint32_t *__restrict__ qnorm, const uint32_t *__restrict__ dnorm, int n_q, int n_dict, int k,
              int32_t *__restrict__ out_idx, uint32_t *__restrict__ out_dist) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sA = smem;
  uint8_t *sB = smem + MT * A_TILE;
  uint8_t *sX = sB + STAGES * B_TILE;                                   // top-k state (TOPK only)
  unsigned long long *s_heap = reinterpret_cast<unsigned long long *>(sX);                  // [KMAX][TK_ROWS]
  unsigned long long *s_queue = s_heap + KMAX * TK_ROWS;                                     // [QCAP][TK_ROWS]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sX + (TOPK ? (KMAX + QCAP) * TK_ROWS * 8 : 0));
  uint64_t *full = bars;                  // [STAGES]  TMA -> MMA
  uint64_t *empty = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t *a_full = bars + 2 * STAGES;   // queries landed
  uint64_t *a_empty = a_full + 1;         // queries no longer read by the tensor pipe
  uint64_t *t_full = a_empty + 1;         // [2] accumulators ready
  uint64_t *t_empty = t_full + 2;         // [2] accumulators drained
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

  constexpr int EPI_WARPS = 4 * MT;
  constexpr int ROWS = BM * MT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (n_dict + BN - 1) / BN;
  const int n_qblocks = (n_q + ROWS - 1) / ROWS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int g = 0; g < 2; ++g) { mbar_init(&t_full[g], 1); mbar_init(&t_empty[g], 4); }
    fence_barrier_init();
  }
  if (warp == EPI_WARPS + 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == EPI_WARPS && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_d); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == EPI_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0, w = 0;
      for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
        mbar_wait(a_empty, (w & 1) ^ 1);
        mbar_expect_tx(a_full, MT * A_TILE);
        for (int g = 0; g < MT; ++g)
          for (int c = 0; c < 3; ++c)
            tma_load_2d(sA + g * A_TILE + c * CHUNK_A, &tmap_q, a_full, c * 128, qb * ROWS + g * BM);
        for (int j = 0; j < n_tiles; ++j, ++it) {
          const uint32_t s = it % STAGES, r = it / STAGES;
          mbar_wait(&empty[s], (r & 1) ^ 1);
          mbar_expect_tx(&full[s], B_TILE);
          for (int c = 0; c < 3; ++c) tma_load_2d(sB + s * B_TILE + c * CHUNK_B, &tmap_d, &full[s], c * 128, j * BN);
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t I_SS = make_idesc(kDFmtS32, kFmtS8, kFmtS8, BM, BN);
      constexpr uint32_t I_SU = make_idesc(kDFmtS32, kFmtS8, kFmtU8, BM, BN);
      constexpr uint32_t I_US = make_idesc(kDFmtS32, kFmtU8, kFmtS8, BM, BN);
      constexpr uint32_t I_UU = make_idesc(kDFmtS32, kFmtU8, kFmtU8, BM, BN);
      const uint64_t descA0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t descB0 = umma_desc_sw128(smem_u32(sB));
      // K-step ks (32 bytes) of a tile lives in chunk ks/4 at byte (ks%4)*32; offsets in 16-byte units
      auto koffA = [](int ks) { return (uint64_t)(((ks >> 2) * CHUNK_A + (ks & 3) * 32) >> 4); };
      auto koffB = [](int ks) { return (uint64_t)(((ks >> 2) * CHUNK_B + (ks & 3) * 32) >> 4); };
      uint32_t it = 0, u = 0, w = 0;
      for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x, ++w) {
        mbar_wait(a_full, w & 1);
        tc_fence_after();
        for (int j = 0; j < n_tiles; ++j, ++it) {
          const uint32_t s = it % STAGES, r = it / STAGES;
          mbar_wait(&full[s], r & 1);
          tc_fence_after();
          const uint64_t dB = descB0 + (uint64_t)((s * B_TILE) >> 4);
          for (int g = 0; g < MT; ++g, ++u) {
            const uint32_t ts = u & 1;
            mbar_wait(&t_empty[ts], ((u >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint64_t dA = descA0 + (uint64_t)((g * A_TILE) >> 4);
            const uint32_t acc = tmem_base + ts * ACC_COLS;
#pragma unroll
            for (int t = 0; t < 6; ++t) mma_i8(acc, dA + koffA(t), dB + koffB(t), I_SS, t > 0);                 // HH
#pragma unroll
            for (int t = 0; t < 6; ++t) mma_i8(acc + BN, dA + koffA(t), dB + koffB(6 + t), I_SU, t > 0);        // HL
#pragma unroll
            for (int t = 0; t < 6; ++t) mma_i8(acc + BN, dA + koffA(6 + t), dB + koffB(t), I_US, 1);            // LH
#pragma unroll
            for (int t = 0; t < 6; ++t) mma_i8(acc + 2 * BN, dA + koffA(6 + t), dB + koffB(6 + t), I_UU, t > 0);  // LL
            tc_commit(&t_full[ts]);
          }
          tc_commit(&empty[s]);
        }
        tc_commit(a_empty);
      }
    }
  } else {
    // ===================== epilogue: thread = query row =====================
    const int g = (MT == 2) ? (warp >> 2) : 0;          // M tile owned by this warp
    const int wq = warp & 3;                            // TMEM lane quarter
    const int row = g * BM + wq * 32 + lane;            // row within the CTA's query block
    const int hrow = wq * 32 + lane;                    // row within the top-k state arrays
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    uint32_t u = (MT == 2) ? (uint32_t)g : 0u;          // this warp's next unit
    for (int qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x) {
      const int64_t qi = (int64_t)qb * ROWS + row;
      const bool valid = qi < n_q;
      const uint32_t nq = valid ? __ldg(qnorm + qi) : 0u;
      uint32_t best_d = 0xFFFFFFFFu;
      int32_t best_i = -1;
      uint32_t tau = 0xFFFFFFFFu;
      int qn = 0;
      if (TOPK) {
        for (int sl = 0; sl < k; ++sl) s_heap[sl * TK_ROWS + hrow] = ~0ull;
      }
      for (int j = 0; j < n_tiles; ++j, u += MT) {
        const uint32_t ts = u & 1;
        mbar_wait(&t_full[ts], (u >> 1) & 1);
        tc_fence_after();
        const uint32_t t_acc = t_lane + ts * ACC_COLS;
        const int col0 = j * BN;
        const int ncol = min(BN, n_dict - col0);
#pragma unroll 1
        for (int ch = 0; ch < BN / 16; ++ch) {
          uint32_t hh[16], xx[16], ll[16];
          tmem_ld16(t_acc + ch * 16, hh);
          tmem_ld16(t_acc + BN + ch * 16, xx);
          tmem_ld16(t_acc + 2 * BN + ch * 16, ll);
          uint32_t nd[16];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 t4 = __ldg(reinterpret_cast<const uint4 *>(dnorm + col0 + ch * 16) + v);
            nd[4 * v] = t4.x; nd[4 * v + 1] = t4.y; nd[4 * v + 2] = t4.z; nd[4 * v + 3] = t4.w;
          }
          tmem_ld_wait();
          if (ch == BN / 16 - 1) {  // accumulators are in registers: hand the TMEM stage back to the tensor pipe
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[ts]);
          }
          // d = nq + nd - 2*(65536*HH + 256*X + LL)  (mod 2^32); m = min over the 16 columns
          uint32_t dv[16];
          uint32_t m = 0xFFFFFFFFu;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            uint32_t d = nq + nd[e];
            d -= hh[e] << 17;
            d -= xx[e] << 9;
            d -= ll[e] << 1;
            dv[e] = d;
            m = min(m, d);
          }
          const int cbase = ch * 16;
          if (ncol < BN) {   // ragged last dictionary tile: columns beyond the dictionary never compete
            m = 0xFFFFFFFFu;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (cbase + e >= ncol) dv[e] = 0xFFFFFFFFu;
              m = min(m, dv[e]);
            }
          }
          if (!TOPK) {
            if (m < best_d) {   // rare once a good candidate has been seen
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (dv[e] < best_d) { best_d = dv[e]; best_i = col0 + cbase + e; }
            }
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t mh = 0xFFFFFFFFu;
#pragma unroll
              for (int e = 0; e < 8; ++e) mh = min(mh, dv[half * 8 + e]);
              if (mh < tau) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const uint32_t d = dv[half * 8 + e];
                  if (d < tau) {
                    s_queue[qn * TK_ROWS + hrow] = ((unsigned long long)d << 32) | (uint32_t)(col0 + cbase + half * 8 + e);
                    ++qn;
                  }
                }
              }
              // a queue may receive 8 more entries before the next check: drain when any row is past half
              if (__any_sync(0xffffffffu, qn > QCAP - 8)) {
                const int qmax = __reduce_max_sync(0xffffffffu, qn);
                for (int t = 0; t < qmax; ++t) {
                  if (t < qn) {
                    const unsigned long long key = s_queue[t * TK_ROWS + hrow];
                    if (key < s_heap[hrow]) heap_replace_root(s_heap, hrow, k, key);
                  }
                }
                qn = 0;
                tau = (uint32_t)(s_heap[hrow] >> 32);
              }
            }
          }
        }
      }
      // ---- results of this query block
      if (!TOPK) {
        if (valid) { out_idx[qi] = best_i; out_dist[qi] = best_d; }
      } else {
        for (int t = 0; t < qn; ++t) {
          const unsigned long long key = s_queue[t * TK_ROWS + hrow];
          if (key < s_heap[hrow]) heap_replace_root(s_heap, hrow, k, key);
        }
        __syncwarp();
        // coalesced write-out: the warp walks its 32 rows, lanes take heap slots
        for (int L = 0; L < 32; ++L) {
          const int64_t qL = (int64_t)qb * ROWS + wq * 32 + L;
          if (qL >= n_q) break;
          for (int p = lane; p < k; p += 32) {
            const unsigned long long key = s_heap[p * TK_ROWS + wq * 32 + L];
            out_idx[qL * k + p] = (int32_t)(uint32_t)key;          // empty slots: 0xFFFFFFFF = -1
            out_dist[qL * k + p] = (uint32_t)(key >> 32);
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) tmem_dealloc(tmem_base, 512);
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

size_t knn_workspace_bytes(int num_ctas) { (void)num_ctas; return 0; }   // top-k state lives in shared memory
int knn_rows_per_cta() { return BM * 2; }

int launch_limb_split(const int16_t *in, int64_t n, uint8_t *limbs, uint32_t *norms, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  limb_split_kernel<<<(unsigned)((n + 7) / 8), 192, 0, st>>>(in, n, limbs, norms);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_knn_i8(const uint8_t *q_limbs, const uint32_t *q_norm, int n_q, const uint8_t *d_limbs, const uint32_t *d_norm,
                  int n_dict, int k, int32_t *out_idx, uint32_t *out_dist, void *ws, int num_ctas, int sort_rows,
                  cudaStream_t st) {
  (void)ws;
  if (n_q <= 0) return TM_OK;
  if (k < 1 || k > KMAX || n_dict <= 0) return TM_ERR_ARG;
  CUtensorMap tq, td;
  int rc = make_tmap_rows_u8(&tq, q_limbs, (uint64_t)n_q, ROWB, BM);
  if (rc != TM_OK) return rc;
  rc = make_tmap_rows_u8(&td, d_limbs, (uint64_t)n_dict, ROWB, BN);
  if (rc != TM_OK) return rc;
  constexpr int SMEM_K1 = 2 * A_TILE + STAGES * B_TILE + 256 + 1024;
  constexpr int SMEM_TK = 1 * A_TILE + STAGES * B_TILE + (KMAX + QCAP) * TK_ROWS * 8 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(knn_i8_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_K1) != cudaSuccess) return TM_ERR_CUDA;
    if (cudaFuncSetAttribute(knn_i8_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TK) != cudaSuccess) return TM_ERR_CUDA;
    attr_set = true;
  }
  {
    ProfScope prof(k == 1 ? "knn_k1" : "knn_topk", st);
    if (k == 1) {
      const int n_qblocks = (n_q + 2 * BM - 1) / (2 * BM);
      const int grid = n_qblocks < num_ctas ? n_qblocks : num_ctas;
      knn_i8_kernel<2, false><<<grid, 64 + 256, SMEM_K1, st>>>(tq, td, q_norm, d_norm, n_q, n_dict, k, out_idx, out_dist);
    } else {
      const int n_qblocks = (n_q + BM - 1) / BM;
      const int grid = n_qblocks < num_ctas ? n_qblocks : num_ctas;
      knn_i8_kernel<1, true><<<grid, 64 + 128, SMEM_TK, st>>>(tq, td, q_norm, d_norm, n_q, n_dict, k, out_idx, out_dist);
    }
  }
  note_launch();
  if (cudaGetLastError() != cudaSuccess) return TM_ERR_CUDA;
  if (k > 1 && sort_rows) {
    note_launch();
    knn_sort_rows_kernel<<<(unsigned)((n_q + 7) / 8), 256, 0, st>>>(out_idx, out_dist, n_q, k);
    if (cudaGetLastError() != cudaSuccess) return TM_ERR_CUDA;
  }
  return TM_OK;
}

}  // namespace tmg
