// motion_tc.cu -- the motion search of TFrame.PredictMotion / TFrame.Reconstruct (tilingencoder.pas:1213-1244, 1496-1532)
// as a dense contraction on the tensor cores.
//
// A CTA takes a block of 8 x 16 frame tiles (128 rows of an MMA) and streams the UNION of their search windows -- at
// most 120 rows x 184 pixel offsets of the previous frame buffer's sliding features -- through the same exact int8-limb
// tcgen05 pipeline as the k-NN kernels (knn_i8.cu): the tile vectors sit in TMEM as the A operand, 64 candidate offsets
// at a time arrive by TMA as the B operand, HH / HL+LH / LL accumulate in TMEM, and eight epilogue warps fold them into
// d = |a|^2 + |b|^2 - 2 a.b (mod 2^32, the reference's Cardinal), add the Manhattan penalty (:1236, :1519), mask the
// offsets outside the row's own 64 x 64 window, and keep the first strict minimum in scan order.  Only 4096 of the
// 22 080 offsets of the union belong to a given tile (18.5 %), but the wasted tensor work is free compared with the
// 272 CUDA-core instructions per useful candidate of the scalar kernel (motion.cu), and every candidate vector is read
// from L2 once per 128 tiles instead of once per tile.
//
// Candidate layout: limb rows [ph][pwp][384] and norms [ph][pwp] with the row pitch pwp = pw rounded up to 64, so that a
// B tile = 64 consecutive offsets of one row starts at a multiple of 64 (TMA box, aligned norm loads).  A thread sees its
// candidates in the reference's row-major scan order, so a strict '<' keeps the first minimum; the two column halves of a
// row are merged on (error, scan index).
#include "tc_common.cuh"
#include "tm_kernels.h"

namespace tmg {

int make_tmap_rows_u8(CUtensorMap *map, const void *base, uint64_t rows, uint32_t row_bytes, uint32_t box_rows);   // knn_i8.cu

namespace {
constexpr int BM = 128, BN = 64, NST = 8, ROWB = 384;
constexpr int CHUNK_B = BN * 128, B_TILE = 3 * CHUNK_B, ACC_COLS = 3 * BN, A_COL = 2 * ACC_COLS;
constexpr int BTW = 16, BTH = 8;   // tiles per block: 16 wide x 8 tall = 128 MMA rows
#ifndef TM_MT_NH
#define TM_MT_NH 4
#endif
// Column splits per B tile = epilogue warps per TMEM lane quarter.  A candidate row is outside the vertical window of two of the
// eight tile rows of a block more often than not, so at any tile ~3 of 8 epilogue warps have nothing to do while the live
// ones set the pace (ncu source page: 40 % of the epilogue warps' samples sit in the accumulator wait); four warps per quarter
// with 16 columns each halve the live warps' work per tile.
constexpr int MT_NH = TM_MT_NH;
constexpr int MT_NEW = 4 * MT_NH;                    // epilogue warps; warp MT_NEW = TMA producer, MT_NEW + 1 / + 2 = MMA issuers
constexpr int MT_THREADS = (MT_NEW + 3) * 32;
}  // namespace

// sliding features [ph * pw][192] int16 -> padded limb rows [ph][pwp][384] + norms [ph][pwp] (pad entries zero)
__global__ void __launch_bounds__(192) cand_limb_split_kernel(const int16_t *__restrict__ in, int pw, int ph, int pwp,
                                                             uint8_t *__restrict__ limbs, uint32_t *__restrict__ norms) {
  __shared__ uint32_t s_norm[8];
  const int r = threadIdx.x / 24, seg = threadIdx.x % 24;
  const int64_t prow = (int64_t)blockIdx.x * 8 + r;          // padded position index
  const int oy = (int)(prow / pwp), ox = (int)(prow - (int64_t)oy * pwp);
  if (threadIdx.x < 8) s_norm[threadIdx.x] = 0;
  __syncthreads();
  const bool real = oy < ph && ox < pw;
  if (oy < ph) {
    uint32_t hi[2] = {0, 0}, lo[2] = {0, 0}, acc = 0;
    if (real) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in + ((int64_t)oy * pw + ox) * 192) + seg);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t a = w[2 * i], b = w[2 * i + 1];
        lo[i] = __byte_perm(a, b, 0x6420);
        hi[i] = __byte_perm(a, b, 0x7531);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int32_t e0 = (int16_t)(w[2 * i + h] & 0xffff), e1 = (int16_t)(w[2 * i + h] >> 16);
          acc += (uint32_t)(e0 * e0) + (uint32_t)(e1 * e1);
        }
      }
    }
    uint8_t *dst = limbs + prow * ROWB + seg * 8;
    *reinterpret_cast<uint2 *>(dst) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2 *>(dst + 192) = make_uint2(lo[0], lo[1]);
    atomicAdd(&s_norm[r], acc);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int64_t pr = (int64_t)blockIdx.x * 8 + threadIdx.x;
    if (pr < (int64_t)ph * pwp) norms[pr] = s_norm[threadIdx.x];
  }
}

#if TM_MT_NH == 4
__global__ void __maxnreg__(96)     // 19 warps: five per register-file partition (16 K registers) -> at most 102 each
#else
__global__ void __launch_bounds__(MT_THREADS, 1)
#endif
motion_tc_kernel(const uint8_t *__restrict__ t_limbs, const uint32_t *__restrict__ t_norm, int tw, int th,
                 const __grid_constant__ CUtensorMap tmap_c, const uint32_t *__restrict__ c_norm, int pwp, int R,
                 int32_t *__restrict__ pred_x, int32_t *__restrict__ pred_y, uint32_t *__restrict__ err_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sB = smem;
  unsigned long long *s_merge = reinterpret_cast<unsigned long long *>(sB + NST * B_TILE);   // [MT_NH - 1][128]
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_merge + BM * (MT_NH - 1));
  uint64_t *full = bars, *empty = bars + NST, *a_full = bars + 2 * NST, *a_empty = a_full + 1, *t_full = a_empty + 1,
           *t_empty = t_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = tw * 8, h = th * 8;
  const int nbx = (tw + BTW - 1) / BTW, nby = (th + BTH - 1) / BTH, n_blocks = nbx * nby;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 4);
    mbar_init(a_empty, 2);
    for (int g = 0; g < 2; ++g) { mbar_init(&t_full[g], 1); mbar_init(&t_empty[g], MT_NEW); }
    fence_barrier_init();
  }
  if (warp == MT_NEW + 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == MT_NEW && lane == 0) tma_prefetch_desc(&tmap_c);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0) __trap();   // the MMA issue code addresses TMEM with immediates

  // geometry of tile block b: union of the windows of its tiles (all roles compute it the same way)
  auto geometry = [&](int b, int &bx0, int &by0, int &OX0, int &OY0, int &n_oy, int &nseg) {
    const int by = b / nbx, bx = b - by * nbx;
    bx0 = bx * BTW; by0 = by * BTH;
    const int tx1 = min(tw, bx0 + BTW) - 1, ty1 = min(th, by0 + BTH) - 1;   // last tile of the block
    const int oxmn = max(0, bx0 * 8 - R - 1), oxmx = min(w - 8, tx1 * 8 + R);
    const int oymn = max(0, by0 * 8 - R - 1), oymx = min(h - 8, ty1 * 8 + R);
    // first candidate column: 16-byte alignment of the norm loads is all that is needed (the TMA box may start at any position;
    // a last segment that runs past the row end reads the next row's positions or the zeroed slack, all masked by the
    // epilogue) -- a 64-aligned start made interior blocks scan 4 segments of 64 columns where their 184 columns need 3
    OX0 = oxmn & ~3; OY0 = oymn;
    n_oy = oymx - oymn + 1;
    nseg = (oxmx - OX0) / 64 + 1;
  };

  if (warp == MT_NEW) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        int bx0, by0, OX0, OY0, n_oy, nseg;
        geometry(b, bx0, by0, OX0, OY0, n_oy, nseg);
        for (int iy = 0; iy < n_oy; ++iy)
          for (int sg = 0; sg < nseg; ++sg, ++it) {
            const uint32_t s = it % NST, r = it / NST;
            mbar_wait(&empty[s], (r & 1) ^ 1);
            mbar_expect_tx(&full[s], B_TILE);
            const int pos = (OY0 + iy) * pwp + OX0 + sg * 64;
            for (int c = 0; c < 3; ++c) tma_load_2d(sB + s * B_TILE + c * CHUNK_B, &tmap_c, &full[s], c * 128, pos);
          }
      }
    }
  } else if (warp > MT_NEW) {
    // ===================== two MMA issuer warps (even / odd B tiles) =====================
    const uint32_t my_parity = (uint32_t)(warp - MT_NEW - 1);
    const uint64_t descB0 = umma_desc_sw128(smem_u32(sB));
    uint32_t it = 0, wv = 0;
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, ++wv) {
      int bx0, by0, OX0, OY0, n_oy, nseg;
      geometry(b, bx0, by0, OX0, OY0, n_oy, nseg);
      const int n_bt = n_oy * nseg;
      mbar_wait(a_full, wv & 1);
      tc_fence_after();
      for (int j = 0; j < n_bt; ++j, ++it) {
        const uint32_t s = it % NST, r = it / NST;
        const uint32_t ts = it & 1;
        if (ts != my_parity) continue;
        mbar_wait(&full[s], r & 1);
        mbar_wait(&t_empty[ts], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dB = descB0 + (uint64_t)((s * B_TILE) >> 4);
#if !(defined(TM_MT_DBG) && (TM_MT_DBG & 2))   // timing experiment: no MMAs
        if (my_parity == 0) mma_i8_tile_elect<0>((uint32_t)dB); else mma_i8_tile_elect<1>((uint32_t)dB);
#endif
        tc_commit_elect(&t_full[ts]);
        tc_commit_elect(&empty[s]);
      }
      tc_commit_elect(a_empty);
    }
  } else {
    // ===================== epilogue: thread = (tile row of the block, column half) =====================
    const int q = warp & 3, hh = warp >> 2;
    const int row = q * 32 + lane;                 // row = ty_local * 16 + tx_local
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int HN = BN / MT_NH;
    uint32_t it = 0, wv = 0;
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x, ++wv) {
      int bx0, by0, OX0, OY0, n_oy, nseg;
      geometry(b, bx0, by0, OX0, OY0, n_oy, nseg);
      const int sx = bx0 + (row & (BTW - 1)), sy = by0 + (row >> 4);
      const bool valid = sx < tw && sy < th;
      const int tile = sy * tw + sx;
      const int dx = sx * 8, dy = sy * 8;
      // this tile's own window (tilingencoder.pas:1213-1216) and the scan-index geometry
      const int oymn = max(0, dy - R - 1), oymx = min(h - 8, dy + R);
      const int oxmn = max(0, dx - R - 1), oxmx = min(w - 8, dx + R);
      const int ww = oxmx - oxmn + 1;
      const uint32_t nq = valid ? __ldg(t_norm + tile) : 0u;
      if (hh == 0) {   // warps 0-3 also store the tile rows into TMEM
        mbar_wait(a_empty, (wv & 1) ^ 1);
        tc_fence_after();
        const uint4 *src = reinterpret_cast<const uint4 *>(t_limbs + (int64_t)(valid ? tile : 0) * ROWB);
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
      uint32_t best = 0xFFFFFFFFu, best_p = 0xFFFFFFFFu;
      // rows of this warp's tiles: two tile rows (32 lanes = 2 x 16 tiles) -> warp-uniform vertical range
      const int wy0 = by0 + 2 * q, wy1 = min(th - 1, wy0 + 1);
      const int w_oymn = max(0, wy0 * 8 - R - 1), w_oymx = wy0 < th ? min(h - 8, wy1 * 8 + R) : -1;
      for (int iy = 0; iy < n_oy; ++iy) {
        const int oy = OY0 + iy;
        const bool warp_row_live = oy >= w_oymn && oy <= w_oymx;
        const bool row_live = valid && oy >= oymn && oy <= oymx;
        const uint32_t pen_y = (uint32_t)abs(oy - dy);
        const uint32_t scan_row = (uint32_t)((oy - oymn) * ww - oxmn);   // scan index = scan_row + ox
        for (int sg = 0; sg < nseg; ++sg, ++it) {
          const uint32_t ts = it & 1;
          const int ox0 = OX0 + sg * 64 + hh * HN;   // first offset of this thread's 32 columns
          // liveness and the candidate norms do not depend on the MMA: decided / fetched BEFORE waiting for the accumulators, so the
          // loads' latency hides behind the wait
          const bool warp_live = warp_row_live && __any_sync(0xffffffffu, row_live && ox0 <= oxmx && ox0 + HN - 1 >= oxmn);
          uint32_t nd[HN];
          if (warp_live) {
            const uint32_t *np = c_norm + (int64_t)oy * pwp + ox0;
#pragma unroll
            for (int v = 0; v < HN / 4; ++v) {
              const uint4 t4 = __ldg(reinterpret_cast<const uint4 *>(np) + v);
              nd[4 * v] = t4.x; nd[4 * v + 1] = t4.y; nd[4 * v + 2] = t4.z; nd[4 * v + 3] = t4.w;
            }
          }
          mbar_wait(&t_full[ts], (it >> 1) & 1);
          tc_fence_after();
          if (!warp_live) {   // no lane of this warp has a candidate in these 32 columns: just hand the stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[ts]);
            continue;
          }
          const uint32_t t_acc = t_lane + ts * ACC_COLS + hh * HN;
          uint32_t pp[HN], xx[HN], lo[HN];
#pragma unroll
          for (int c = 0; c < HN / 16; ++c) {
            tmem_ld16(t_acc + c * 16, reinterpret_cast<uint32_t(&)[16]>(pp[c * 16]));
            tmem_ld16(t_acc + BN + c * 16, reinterpret_cast<uint32_t(&)[16]>(xx[c * 16]));
            tmem_ld16(t_acc + 2 * BN + c * 16, reinterpret_cast<uint32_t(&)[16]>(lo[c * 16]));
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[ts]);
#if defined(TM_MT_DBG) && (TM_MT_DBG & 1)
          continue;   // timing experiment: no epilogue arithmetic
#endif
          if (!row_live) continue;
          const int e_lo = max(0, oxmn - ox0), e_hi = min(HN - 1, oxmx - ox0);   // this row's live columns
          if (e_lo > e_hi) continue;
          const uint32_t live_cols = ((2u << e_hi) - 1u) & ~((1u << e_lo) - 1u);   // bits e_lo..e_hi
          uint32_t m = 0xFFFFFFFFu;
          const uint32_t base = nq + pen_y;
          const int v0 = ox0 - dx;
#pragma unroll
          for (int e = 0; e < HN; ++e) {   // error incl. the Manhattan penalty, kept in pp[]; columns outside the window never compete
            // d = |a|^2 + |b|^2 + penalty - 2 (65536 HH + 256 X + LL)  (mod 2^32): (256 HH + X) * (-512) + s, then - LL - LL
            const uint32_t s3 = nd[e] + base + (uint32_t)abs(v0 + e);
            const uint32_t a = pp[e] * 256u + xx[e];
            uint32_t d = a * 0xFFFFFE00u + s3;
            d = d - lo[e] - lo[e];
            d = (live_cols & (1u << e)) ? d : 0xFFFFFFFFu;   // one bit test per column instead of two range compares
            pp[e] = d;
            m = min(m, d);
          }
          if (m < best) {   // strict: an equal error later in the scan never replaces the first minimum
#pragma unroll
            for (int e = HN - 1; e >= 0; --e)
              if (pp[e] == m) best_p = scan_row + (uint32_t)(ox0 + e);   // lowest e wins
            best = m;
          }
        }
      }
      // merge the two column halves of each row on (error, scan index)
      const unsigned long long key = ((unsigned long long)best << 32) | best_p;
      if (hh > 0) s_merge[(hh - 1) * BM + row] = key;
      asm volatile("bar.sync 1, %0;\n" ::"n"(MT_NEW * 32) : "memory");
      if (hh == 0 && valid) {
        unsigned long long k2 = key;
#pragma unroll
        for (int o2 = 0; o2 < MT_NH - 1; ++o2) { const unsigned long long o = s_merge[o2 * BM + row]; k2 = o < k2 ? o : k2; }
        const uint32_t e = (uint32_t)(k2 >> 32), p = (uint32_t)k2;
        int bx = 0, byy = 0;
        if (e != 0xFFFFFFFFu || p != 0xFFFFFFFFu) {
          const int wy = (int)(p / (uint32_t)ww), wx = (int)(p - (uint32_t)wy * (uint32_t)ww);
          bx = oxmn + wx - dx; byy = oymn + wy - dy;
        }
        pred_x[tile] = bx; pred_y[tile] = byy; err_out[tile] = e;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(MT_NEW * 32) : "memory");   // s_merge is rewritten by the next tile block
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MT_NEW + 1) tmem_dealloc(tmem_base, 512);
}

size_t motion_tc_ws_bytes(int tw, int th) {
  const int pw = tw * 8 - 7, ph = th * 8 - 7, pwp = (pw + 63) & ~63;
  const size_t npad = (size_t)ph * pwp + 256;   // + slack: the last B tile of the last row may read 64 positions past the end
  return npad * ROWB + npad * 4 + (size_t)tw * th * (ROWB + 4) + 4096;
}

// where the candidate operands of a search live inside its workspace (for a producer that writes them directly:
// launch_features_sliding_limbs); launch_motion_search_tc(dcts = nullptr) then skips its own split pass
void motion_tc_cand_layout(void *ws, int tw, int th, uint8_t **c_limbs, uint32_t **c_norm, int *pwp_out) {
  const int pw = tw * 8 - 7, ph = th * 8 - 7, pwp = (pw + 63) & ~63;
  const size_t npad = (size_t)ph * pwp + 256;
  *c_limbs = (uint8_t *)ws;
  *c_norm = (uint32_t *)((uint8_t *)ws + npad * ROWB);
  *pwp_out = pwp;
}

int launch_motion_search_tc(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting, int32_t *pred_x,
                            int32_t *pred_y, uint32_t *err, void *ws, size_t ws_bytes, int num_ctas, cudaStream_t st) {
  if (tw < 1 || th < 1 || radius_setting < 1) return TM_ERR_ARG;
  if (ws_bytes < motion_tc_ws_bytes(tw, th)) return TM_ERR_ARG;
  const int pw = tw * 8 - 7, ph = th * 8 - 7, pwp = (pw + 63) & ~63;
  const size_t npad = (size_t)ph * pwp + 256;
  uint8_t *p = (uint8_t *)ws;
  uint8_t *c_limbs = p; p += npad * ROWB;
  uint32_t *c_norm = (uint32_t *)p; p += npad * 4;
  p = (uint8_t *)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  uint8_t *t_limbs = p; p += (size_t)tw * th * ROWB;
  uint32_t *t_norm = (uint32_t *)p;
  ProfScope prof("motion_search", st);
  const int64_t nprow = (int64_t)ph * pwp;
  if (dcts) cand_limb_split_kernel<<<(unsigned)((nprow + 7) / 8), 192, 0, st>>>(dcts, pw, ph, pwp, c_limbs, c_norm);
  if (cudaMemsetAsync(c_limbs + (size_t)nprow * ROWB, 0, 256 * ROWB, st) != cudaSuccess) return TM_ERR_CUDA;
  if (cudaMemsetAsync(c_norm + nprow, 0, 256 * 4, st) != cudaSuccess) return TM_ERR_CUDA;
  int rc = launch_limb_split(cur_feat, (int64_t)tw * th, t_limbs, t_norm, st);
  if (rc) return rc;
  CUtensorMap tc;
  rc = make_tmap_rows_u8(&tc, c_limbs, (uint64_t)npad, ROWB, BN);
  if (rc != TM_OK) return rc;
  constexpr int SMEM = NST * B_TILE + BM * 8 * (MT_NH - 1) + 256 + 1024;
  static bool attr_set[TM_MAX_DEVICES] = {};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(motion_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return TM_ERR_CUDA;
  }
  const int n_blocks = ((tw + BTW - 1) / BTW) * ((th + BTH - 1) / BTH);
  const int grid = n_blocks < num_ctas ? n_blocks : num_ctas;
  motion_tc_kernel<<<grid, MT_THREADS, SMEM, st>>>(t_limbs, t_norm, tw, th, tc, c_norm, pwp, radius_setting - 1, pred_x, pred_y, err);
  note_launch(3);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
