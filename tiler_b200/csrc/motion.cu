// motion.cu -- the per-frame motion search and the Reconstruct decision of TFrame.Reconstruct / TFrame.PredictMotion
// (tilingencoder.pas:1154-1282, 1430-1679), device-resident so that the frame -> frame chain of a
// keyframe sequence never leaves the GPU.
//
//   motion_search_kernel : one CTA per frame tile.  The tile's window of <= 64 x 64 pixel offsets is scanned by 256
//       threads; every candidate is a 192-d int16 vector of the previous frame buffer's sliding features, the error is
//       sum (a-b)^2 (uint32, wraps like the reference's Cardinal) + the Manhattan penalty (:1236, :1519).  The
//       reference keeps the FIRST strict minimum of a row-major scan: every thread scans its offsets in increasing
//       scan order, the block reduces on (error, scan index).  QuickTest (utils.pas:755-759) is an optimisation only
//       (the sum over the first 8 coefficients bounds the full error from below); here it prunes against the best
//       error found so far by ANY warp of the block (shared-memory atomicMin), with a STRICT comparison so that a
//       candidate that could still tie -- and win on scan order -- is always evaluated: the result does not depend on
//       thread timing.  A pruned candidate costs one 32-byte sector instead of twelve.
//   reconstruct_decide_kernel : one 64-thread group per tile: dead band on the motion error (:1534), KNN vs motion with the
//       192 tolerance (:1614), draw into the front buffer (:1623-1651), error -> PSNR (utils.pas:1074-1078).
#include "tm_kernels.h"
#include <math.h>

namespace tmg {

__device__ __forceinline__ uint32_t sq_diff8(const uint4 a, const uint4 b, uint32_t acc) {
  const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int32_t d0 = (int32_t)(int16_t)(wa[i] & 0xffff) - (int32_t)(int16_t)(wb[i] & 0xffff);
    const int32_t d1 = ((int32_t)wa[i] >> 16) - ((int32_t)wb[i] >> 16);
    acc += (uint32_t)(d0 * d0);
    acc += (uint32_t)(d1 * d1);
  }
  return acc;
}

// Warp-cooperative scan.  A warp takes 32 consecutive candidates of the row-major scan at a time:
//   phase A  lane l reads the first 8 coefficients (one 32-byte sector) of candidate base + l and applies QuickTest
//            against the block-wide best error (strict, so a candidate that could tie is never dropped);
//   phase B  the survivors are evaluated four at a time in scan order, one per 8-lane group: a lane loads three 16-byte
//            chunks (each group load is one 128-byte line), square-differences them against its own three chunks of the
//            tile's vector, and three shuffles sum the group.  The tile's vector costs 16 registers per lane, not 96.
// The zero-motion candidate is evaluated first to seed the block-wide best: it is a real candidate, so pruning against
// its error is exact, and on real clips it removes most of the window after 8 coefficients.
__global__ void __launch_bounds__(256) motion_search_kernel(const int16_t *__restrict__ cur_feat, int tw, int th,
                                                            const int16_t *__restrict__ dcts, int R, int32_t *__restrict__ pred_x,
                                                            int32_t *__restrict__ pred_y, uint32_t *__restrict__ err_out) {
  __shared__ unsigned long long s_best[8];
  __shared__ uint32_t s_min;
  const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sy = t / tw, sx = t - sy * tw, dx = sx * 8, dy = sy * 8;
  const int w = tw * 8, h = th * 8, pw = w - 7;
  const uint4 *curp = reinterpret_cast<const uint4 *>(cur_feat + (int64_t)t * 192);
  const int g = lane >> 3, j = lane & 7;            // phase B: 8 lanes per candidate, 3 chunks per lane
  const uint4 cur0 = __ldg(curp);
  const uint4 cj0 = __ldg(curp + j), cj1 = __ldg(curp + j + 8), cj2 = __ldg(curp + j + 16);
  const int oymn = max(0, dy - R - 1), oymx = min(h - 8, dy + R);
  const int oxmn = max(0, dx - R - 1), oxmx = min(w - 8, dx + R);
  const int ww = oxmx - oxmn + 1, wh = oymx - oymn + 1, np = ww * wh;
  // full error of scan position q (< 0: none) computed by the 8 lanes of a group; every lane of the group gets the sum
  auto group_error = [&](int q) -> uint32_t {
    uint32_t e = 0;
    int oy = 0, ox = 0;
    if (q >= 0) {
      const int wy = q / ww, wx = q - wy * ww;
      oy = oymn + wy; ox = oxmn + wx;
      const uint4 *src = reinterpret_cast<const uint4 *>(dcts + ((int64_t)oy * pw + ox) * 192);
      const uint4 v0 = __ldg(src + j), v1 = __ldg(src + j + 8), v2 = __ldg(src + j + 16);   // 8 lanes x 16 B = one 128-byte line each
      e = sq_diff8(cj0, v0, 0u);
      e = sq_diff8(cj1, v1, e);
      e = sq_diff8(cj2, v2, e);
    }
    e += __shfl_xor_sync(0xffffffffu, e, 4);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    return q >= 0 ? e + (uint32_t)(abs(ox - dx) + abs(oy - dy)) : 0xFFFFFFFFu;
  };
  if (warp == 0) {   // seed: the zero-motion candidate
    const int qc = (dy - oymn) * ww + (dx - oxmn);
    const uint32_t e = group_error(g == 0 ? qc : -1);
    if (lane == 0) s_min = e;
  }
  __syncthreads();
  uint32_t best = 0xFFFFFFFFu, best_p = 0xFFFFFFFFu;   // warp-uniform
  for (int base = warp * 32; base < np; base += 256) {
    const int p = base + lane;
    bool alive = false;
    if (p < np) {
      const int wy = p / ww, wx = p - wy * ww;
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(dcts + ((int64_t)(oymn + wy) * pw + oxmn + wx) * 192));
      alive = sq_diff8(cur0, v, 0u) <= *(volatile uint32_t *)&s_min;   // QuickTestEuclideanDCTPtr (strict prune: ties are evaluated)
    }
    uint32_t m = __ballot_sync(0xffffffffu, alive);
    while (m) {   // four survivors per pass, one per 8-lane group, taken in scan order
      const int cnt = __popc(m);
      const int b = g < cnt ? (int)__fns(m, 0, g + 1) : -1;
      const int q = b >= 0 ? base + b : -1;
      const uint32_t e = group_error(q);
      bool improved = false;
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        const uint32_t eg = __shfl_sync(0xffffffffu, e, gg * 8);
        const int qg = __shfl_sync(0xffffffffu, q, gg * 8);
        if (qg >= 0 && eg < best) { best = eg; best_p = (uint32_t)qg; improved = true; }
      }
      if (improved && lane == 0) atomicMin(&s_min, best);
#pragma unroll
      for (int k = 0; k < 4; ++k) m &= m - 1;   // m - 1 on 0 wraps, and 0 & x = 0
    }
  }
  // block arg-min on (error, scan index): the first strict minimum of the row-major scan
  if (lane == 0) s_best[warp] = ((unsigned long long)best << 32) | best_p;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long key = s_best[0];
    for (int i = 1; i < 8; ++i) key = s_best[i] < key ? s_best[i] : key;
    const uint32_t e = (uint32_t)(key >> 32), p = (uint32_t)key;
    int bx = 0, by = 0;
    if (e != 0xFFFFFFFFu && p != 0xFFFFFFFFu) { const int wy = (int)p / ww, wx = (int)p - wy * ww; bx = oxmn + wx - dx; by = oymn + wy - dy; }
    pred_x[t] = bx; pred_y[t] = by; err_out[t] = e;
  }
}

int launch_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting, int32_t *pred_x,
                         int32_t *pred_y, uint32_t *err, cudaStream_t st) {
  if (tw < 1 || th < 1 || radius_setting < 1) return TM_ERR_ARG;
  ProfScope prof("motion_search", st);
  motion_search_kernel<<<tw * th, 256, 0, st>>>(cur_feat, tw, th, dcts, radius_setting - 1, pred_x, pred_y, err);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// EuclideanToPSNR, utils.pas:1074-1078: Single(d / 192), max 0.5, 10 log10(255^2 / x) -> Single
__device__ __forceinline__ float euclidean_to_psnr(uint32_t d) {
  const float r = (float)((double)d * (1.0 / 192.0));
  const double m = r > 0.5f ? (double)r : 0.5;
  return (float)(10.0 * log10(255.0 * 255.0 / m));
}

__global__ void __launch_bounds__(256) reconstruct_decide_kernel(
    const uint8_t *__restrict__ flags, int tw, int th, const int32_t *__restrict__ mp_x, const int32_t *__restrict__ mp_y,
    const uint32_t *__restrict__ mp_err, const int32_t *__restrict__ knn_tile, const int32_t *__restrict__ knn_pal,
    const uint32_t *__restrict__ knn_err, const uint8_t *__restrict__ dict_idx, const int32_t *__restrict__ palettes, int pal_size,
    const int32_t *__restrict__ back, int32_t *__restrict__ front, int32_t *__restrict__ tile_idx, int32_t *__restrict__ pal_idx,
    int32_t *__restrict__ pred_x, int32_t *__restrict__ pred_y, uint8_t *__restrict__ is_pred, uint32_t *__restrict__ err_out,
    float *__restrict__ psnr_out) {
  const int t = blockIdx.x * 4 + (threadIdx.x >> 6), px = threadIdx.x & 63;
  if (t >= tw * th) return;
  const int sy = t / tw, sx = t - sy * tw, dx = sx * 8, dy = sy * 8, w = tw * 8;
  const bool motion = mp_err != nullptr;
  const uint32_t mp = motion ? mp_err[t] : 0xFFFFFFFFu;
  int32_t ti, pi; uint32_t ke;
  if (mp <= 192u) { ti = -1; pi = -1; ke = 0xFFFFFFFFu; }           // IsZero(mpErr, cTileDCTSize)
  else { ti = knn_tile[t]; pi = knn_pal[t]; ke = knn_err[t]; if (ti < 0) { pi = -1; ke = 0xFFFFFFFFu; } }
  const bool knn_best = (unsigned long long)ke + 192ull < (unsigned long long)mp;   // CompareValue(knnErr, mpErr, 192) < 0
  const int mx = motion ? mp_x[t] : 0, my = motion ? mp_y[t] : 0;
  const int ty = px >> 3, tx = px & 7;
  int32_t col = 0;
  if (knn_best) {
    const int fl = flags[t];
    const int tym = (fl & 2) ? 7 - ty : ty, txm = (fl & 1) ? 7 - tx : tx;
    col = __ldg(palettes + (int64_t)pi * pal_size + __ldg(dict_idx + (int64_t)ti * 64 + tym * 8 + txm));
  } else if (motion) {
    col = back[(int64_t)(dy + ty + my) * w + dx + tx + mx];
  }
  front[(int64_t)(dy + ty) * w + dx + tx] = col;
  if (px == 0) {
    const uint32_t e = knn_best ? ke : mp;
    tile_idx[t] = ti; pal_idx[t] = pi; pred_x[t] = mx; pred_y[t] = my; is_pred[t] = knn_best ? 0 : 1; err_out[t] = e;
    if (psnr_out) psnr_out[t] = euclidean_to_psnr(e);
  }
}

int launch_reconstruct_decide(const uint8_t *flags, int tw, int th, const int32_t *mp_x, const int32_t *mp_y, const uint32_t *mp_err,
                              const int32_t *knn_tile, const int32_t *knn_pal, const uint32_t *knn_err, const uint8_t *dict_idx,
                              const int32_t *palettes, int pal_size, const int32_t *back, int32_t *front, int32_t *tile_idx,
                              int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred, uint32_t *err, float *psnr,
                              cudaStream_t st) {
  const int nt = tw * th;
  reconstruct_decide_kernel<<<(nt + 3) / 4, 256, 0, st>>>(flags, tw, th, mp_x, mp_y, mp_err, knn_tile, knn_pal, knn_err, dict_idx, palettes,
                                                         pal_size, back, front, tile_idx, pal_idx, pred_x, pred_y, is_pred, err, psnr);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

__global__ void __launch_bounds__(256) sq_err_rgb_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n,
                                                         unsigned long long *__restrict__ acc) {
  unsigned long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t x = a[i], y = b[i];
#pragma unroll
    for (int sh = 0; sh < 24; sh += 8) { const int d = ((x >> sh) & 255) - ((y >> sh) & 255); s += (unsigned long long)(d * d); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(acc, s);
}

int launch_sq_err_rgb(const int32_t *a, const int32_t *b, int64_t n, unsigned long long *acc, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  sq_err_rgb_kernel<<<592, 256, 0, st>>>(a, b, n, acc);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
