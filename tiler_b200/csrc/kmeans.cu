// kmeans.cu -- Lloyd's algorithm with yakmo's contract (extern.pas:198-203; call sites tilingencoder.pas:4198-4207 for
// the 192-d palette clustering and :4492-4500 for 3-d colour quantisation), plus the colour-quantisation tail of
// QuantizeUsingYakmo / DoQuantization (:4511-4564).
//
//   f64 path (any dim): assignment with the oracle's summation order (bit-identical distances, first minimum), update
//     as an ORDERED segmented sum -- points are stably sorted by label (cub radix sort) and one thread per
//     (cluster, dimension) adds its members in index order, so centroids are bit-identical to a sequential CPU loop
//     and independent of scheduling.
//   RGB path (dim 3, integer points): sums are integers < 2^53, so any summation order is exact; per-block shared
//     memory accumulators + a few global integer atomics per block.  All palettes are clustered in one launch per
//     Lloyd iteration (pixels grouped by palette).
#include "tm_kernels.h"
#include <cub/cub.cuh>
#include <math.h>
#include <vector>

namespace tmg {

__device__ __forceinline__ unsigned long long xorshift64s(unsigned long long &s) {
  unsigned long long x = s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  s = x;
  return x * 0x2545F4914F6CDD1DULL;
}

// ------------------------------------------------------------------ f64 assignment
constexpr int KA_P = 32;  // points per block (lane = point), 4 warps split the centroids
__global__ void __launch_bounds__(128)
kmeans_assign_f64_kernel(const double *__restrict__ x, int64_t n, int dim, const double *__restrict__ cent, int k,
                         int32_t *__restrict__ labels, double *__restrict__ dist, int32_t *__restrict__ changed) {
  extern __shared__ double s_x[];  // [KA_P][dim+1]
  __shared__ double s_best[4][KA_P];
  __shared__ int32_t s_bi[4][KA_P];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p0 = (int64_t)blockIdx.x * KA_P;
  const int ld = dim + 1;
  for (int i = threadIdx.x; i < KA_P * dim; i += 128) {
    const int r = i / dim, c = i % dim;
    s_x[r * ld + c] = (p0 + r < n) ? x[(p0 + r) * dim + c] : 0.0;
  }
  __syncthreads();
  const double *xv = s_x + lane * ld;
  const int per = (k + 3) / 4, lo = w * per, hi = min(lo + per, k);
  double best = INFINITY;
  int32_t bi = -1;
  for (int c = lo; c < hi; ++c) {
    const double *cv = cent + (int64_t)c * dim;
    double s = 0.0;
    for (int j = 0; j < dim; ++j) {
      const double df = __dsub_rn(xv[j], __ldg(cv + j));
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    if (s < best) { best = s; bi = c; }  // NaN centroids (empty clusters) never win
  }
  s_best[w][lane] = best;
  s_bi[w][lane] = bi;
  __syncthreads();
  if (w == 0 && p0 + lane < n) {
    for (int o = 1; o < 4; ++o)
      if (s_best[o][lane] < best) { best = s_best[o][lane]; bi = s_bi[o][lane]; }
    const int32_t old = labels[p0 + lane];
    if (bi < 0) bi = old >= 0 ? old : 0;
    if (bi != old) { labels[p0 + lane] = bi; atomicAdd(changed, 1); }
    if (dist) dist[p0 + lane] = best;
  }
}

// ------------------------------------------------------------------ f64 ordered update
__global__ void iota_kernel(int32_t *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}
__global__ void hist_kernel(const int32_t *__restrict__ labels, int64_t n, int32_t *__restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(counts + labels[i], 1);
}
// Per-cluster sums in the oracle's FIXED blocked order (tmo_kmeans_lloyd): members in ascending point index, consecutive
// blocks of KM_BLOCK members summed sequentially from 0, block sums added sequentially from 0.  A cluster of <= KM_BLOCK
// members is the plain sequential sum.  Members of a cluster are contiguous in `order` (stable sort by label).
constexpr int KM_BLOCK = 128;

// blocks of the clusters with more than KM_BLOCK members (they are the only ones that need partial sums)
__global__ void kmeans_nblk_kernel(const int32_t *__restrict__ counts, int k, int32_t *__restrict__ nb) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < k) nb[c] = counts[c] > KM_BLOCK ? (counts[c] + KM_BLOCK - 1) / KM_BLOCK : 0;
}

// CTA = one member block of one large cluster; thread = dimension (strided).  partial[b][j] (and pw[b] for weights)
template <typename T>
__global__ void __launch_bounds__(256)
kmeans_partial_kernel(const T *__restrict__ x, const double *__restrict__ wts, int dim, const int32_t *__restrict__ order,
                      const int32_t *__restrict__ offs, const int32_t *__restrict__ counts, const int32_t *__restrict__ nb,
                      const int32_t *__restrict__ poffs, int k, double *__restrict__ partial, double *__restrict__ pw) {
  const int b = blockIdx.x;
  // last cluster c with poffs[c] <= b: zero-width (small) clusters share their successor's offset, so this is the owner
  int lo = 0, hi = k - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (poffs[mid] <= b) lo = mid; else hi = mid - 1;
  }
  const int c = lo;
  const int rel = b - poffs[c];
  if (rel >= nb[c]) return;   // beyond the last block (the grid is an upper bound)
  const int cnt = counts[c], o0 = offs[c] + rel * KM_BLOCK;
  const int m1 = min(KM_BLOCK, cnt - rel * KM_BLOCK);
  __shared__ int32_t s_idx[KM_BLOCK];
  for (int m = threadIdx.x; m < m1; m += blockDim.x) s_idx[m] = order[o0 + m];
  __syncthreads();
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    double s = 0.0;
    if (wts) {
      for (int m = 0; m < m1; ++m) {
        const int32_t i = s_idx[m];
        s = __dadd_rn(s, __dmul_rn(__ldg(wts + i), (double)__ldg(x + (int64_t)i * dim + j)));
      }
    } else {
      for (int m = 0; m < m1; ++m) s = __dadd_rn(s, (double)__ldg(x + (int64_t)s_idx[m] * dim + j));
    }
    partial[(int64_t)b * dim + j] = s;
  }
  if (wts && threadIdx.x == 0) {
    double ws = 0.0;
    for (int m = 0; m < m1; ++m) ws = __dadd_rn(ws, __ldg(wts + s_idx[m]));
    pw[b] = ws;
  }
}

// thread = (cluster, dim): small clusters are summed here, large ones add up their block sums
template <typename T>
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const T *__restrict__ x, const double *__restrict__ wts, int dim, const int32_t *__restrict__ order,
                         const int32_t *__restrict__ offs, const int32_t *__restrict__ counts, const int32_t *__restrict__ nb,
                         const int32_t *__restrict__ poffs, const double *__restrict__ partial, const double *__restrict__ pw,
                         int k, int nan_empty, int divide,
                         double *__restrict__ cent, int64_t *__restrict__ counts_out, double *__restrict__ wsum_out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)k * dim) return;
  const int c = (int)(t / dim), j = (int)(t % dim);
  const int cnt = counts[c], o0 = offs[c];
  if (j == 0 && counts_out) counts_out[c] = cnt;
  if (cnt == 0) {
    if (!divide) cent[t] = 0.0;
    else if (nan_empty) cent[t] = NAN;
    if (j == 0 && wsum_out) wsum_out[c] = 0.0;
    return;
  }
  double s = 0.0, ws = 0.0;
  if (cnt > KM_BLOCK) {
    const int b0 = poffs[c], n_b = nb[c];
    for (int b = 0; b < n_b; ++b) s = __dadd_rn(s, partial[(int64_t)(b0 + b) * dim + j]);
    if (wts) { for (int b = 0; b < n_b; ++b) ws = __dadd_rn(ws, pw[b0 + b]); }
    else ws = (double)cnt;
  } else if (wts) {
    for (int m = 0; m < cnt; ++m) {
      const int32_t i = order[o0 + m];
      const double w = __ldg(wts + i);
      s = __dadd_rn(s, __dmul_rn(w, (double)__ldg(x + (int64_t)i * dim + j)));
      ws = __dadd_rn(ws, w);
    }
  } else {
    for (int m = 0; m < cnt; ++m) s = __dadd_rn(s, (double)__ldg(x + (int64_t)order[o0 + m] * dim + j));
    ws = (double)cnt;
  }
  cent[t] = divide ? __ddiv_rn(s, ws) : s;
  if (j == 0 && wsum_out) wsum_out[c] = ws;
}

__global__ void kmeans_finish_kernel(const double *__restrict__ sums, const int64_t *__restrict__ counts, int k, int dim, int nan_empty,
                                     double *__restrict__ cent) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)k * dim) return;
  const int64_t cnt = counts[t / dim];
  if (cnt > 0) cent[t] = __ddiv_rn(sums[t], (double)cnt);
  else if (nan_empty) cent[t] = NAN;
}

// k-means++ seeding for f64 rows, same draw sequence as the oracle's tmo_kmeanspp_init (xorshift64*, D^2 sampling by the
// first prefix sum exceeding u).  Per centroid: a grid-wide kernel refreshes every point's squared distance to the last
// centroid (one thread per point, the oracle's sequential summation order over the dimensions), then ONE block draws the
// next centroid: 256 threads sum contiguous slices of d2 in index order, thread 0 adds the 256 partials in order, draws u
// and walks the owning slice.  (The first version did the distances in that single block as well: 110 ms for 83 000 x 192
// points and 16 centroids, half of PreparePalettes.)
__global__ void __launch_bounds__(128)
kmpp_dist_kernel(const double *__restrict__ x, int64_t n, int dim, const double *__restrict__ last, int first, double *__restrict__ d2) {
  extern __shared__ double s_last[];  // [dim]
  for (int j = threadIdx.x; j < dim; j += 128) s_last[j] = last[j];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  const double *xi = x + i * dim;
  double sum = 0.0;
  for (int j = 0; j < dim; ++j) { const double df = __dsub_rn(__ldg(xi + j), s_last[j]); sum = __dadd_rn(sum, __dmul_rn(df, df)); }
  if (first || sum < d2[i]) d2[i] = sum;
}

// state[0] = xorshift state.  c == 0: draw the first centroid uniformly; else D^2-sample centroid c from d2.
__global__ void __launch_bounds__(256)
kmpp_pick_kernel(const double *__restrict__ x, int64_t n, int dim, int c, unsigned long long seed, const double *__restrict__ d2,
                 unsigned long long *__restrict__ state, double *__restrict__ cent) {
  __shared__ double s_part[256];
  __shared__ long long s_pick;
  const int t = threadIdx.x;
  const int64_t per = (n + 255) / 256, a = min((int64_t)t * per, n), b = min(a + per, n);
  if (c == 0) {
    if (t == 0) {
      unsigned long long st = seed ? seed : 0x9E3779B97F4A7C15ULL;
      s_pick = (long long)(xorshift64s(st) % (unsigned long long)n);
      state[0] = st;
    }
  } else {
    double part = 0.0;
    for (int64_t i = a; i < b; ++i) part = __dadd_rn(part, d2[i]);
    s_part[t] = part;
    __syncthreads();
    if (t == 0) {
      unsigned long long st = state[0];
      double total = 0.0;
      for (int i = 0; i < 256; ++i) total = __dadd_rn(total, s_part[i]);
      const double u = (double)(xorshift64s(st) >> 11) * (1.0 / 9007199254740992.0) * total;
      state[0] = st;
      double acc = 0.0;
      int owner = 255;
      for (int i = 0; i < 256; ++i) {
        if (__dadd_rn(acc, s_part[i]) > u) { owner = i; break; }
        acc = __dadd_rn(acc, s_part[i]);
      }
      const int64_t oa = min((int64_t)owner * per, n), ob = min(oa + per, n);
      long long pk = n - 1;
      for (int64_t i = oa; i < ob; ++i) {
        acc = __dadd_rn(acc, d2[i]);
        if (acc > u) { pk = i; break; }
      }
      s_pick = pk;
    }
  }
  __syncthreads();
  const int64_t pick = s_pick;
  for (int j = t; j < dim; j += 256) cent[(int64_t)c * dim + j] = x[pick * dim + j];
}

static inline size_t km_al(size_t v) { return (v + 255) & ~(size_t)255; }
// member blocks of large clusters: each has > KM_BLOCK members, so there are at most 2n / KM_BLOCK of them
static inline int64_t km_max_blocks(int64_t n) { return 2 * (n / KM_BLOCK) + 1; }
// stats[0] += labels changed in this step (device counter), stats[1] += points that needed the exact f64 scan
__global__ void kmeans_stats_add_kernel(long long *__restrict__ stats, const int32_t *__restrict__ counters, long long n_bf) {
  stats[0] += (long long)counters[0];
  stats[1] += n_bf;
}
int launch_kmeans_stats_add(int64_t *stats, const int32_t *counters, int64_t n_bf, cudaStream_t st) {
  kmeans_stats_add_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long *>(stats), counters, (long long)n_bf);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}
// out[0] = sum of v[0..n) (reporting scalar: block partials in index order, then one thread adds the partials in order)
__global__ void __launch_bounds__(256) sum_f64_partial_kernel(const double *__restrict__ v, int64_t n, double *__restrict__ part) {
  __shared__ double s[256];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  double a = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) a += v[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
__global__ void sum_f64_final_kernel(const double *__restrict__ part, int nb, double *__restrict__ out) {
  double a = 0.0;
  for (int i = 0; i < nb; ++i) a += part[i];
  out[0] = a;
}
int launch_sum_f64(const double *v, int64_t n, double *part /* >= 1024 doubles */, double *out, cudaStream_t st) {
  const int nb = (int)(n < 1024 * 256 ? (n + 255) / 256 : 1024);
  sum_f64_partial_kernel<<<nb > 0 ? nb : 1, 256, 0, st>>>(v, n, part);
  sum_f64_final_kernel<<<1, 1, 0, st>>>(part, nb > 0 ? nb : 1, out);
  note_launch(2);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

size_t kmeans_update_ws_bytes(int64_t n, int k, int dim) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, k);
  size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  return km_al(tmp) + 3 * km_al((size_t)n * 4) + 4 * km_al((size_t)k * 4) + km_al((size_t)km_max_blocks(n) * dim * 8) +
         km_al((size_t)km_max_blocks(n) * 8);
}

int launch_kmeans_assign_f64(const double *x, int64_t n, int dim, const double *cent, int k, int32_t *labels, double *dist,
                             int32_t *changed, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  if (dim < 1 || dim > 1024 || k < 1) return TM_ERR_ARG;
  const size_t smem = (size_t)KA_P * (dim + 1) * sizeof(double);
  if (smem > 48 * 1024)
    if (cudaFuncSetAttribute(kmeans_assign_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return TM_ERR_CUDA;
  kmeans_assign_f64_kernel<<<(unsigned)((n + KA_P - 1) / KA_P), 128, smem, st>>>(x, n, dim, cent, k, labels, dist, changed);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

template <typename T>
static int launch_kmeans_update_t(const T *x, const double *weights, int64_t n, int dim, const int32_t *labels, int k, double *cent,
                                  int64_t *counts_out, double *wsum_out, void *ws, size_t ws_bytes, int nan_empty, int divide,
                                  cudaStream_t st) {
  if (n <= 0) return TM_OK;
  if (ws_bytes < kmeans_update_ws_bytes(n, k, dim)) return TM_ERR_ARG;
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                  (int32_t *)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, k);
  size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  const int64_t max_blocks = km_max_blocks(n);
  uint8_t *p = (uint8_t *)ws;
  void *d_tmp = p; p += km_al(tmp);
  int32_t *keys_out = (int32_t *)p; p += km_al((size_t)n * 4);
  int32_t *vals_in = (int32_t *)p; p += km_al((size_t)n * 4);
  int32_t *vals_out = (int32_t *)p; p += km_al((size_t)n * 4);
  int32_t *counts = (int32_t *)p; p += km_al((size_t)k * 4);
  int32_t *offs = (int32_t *)p; p += km_al((size_t)k * 4);
  int32_t *nb = (int32_t *)p; p += km_al((size_t)k * 4);
  int32_t *poffs = (int32_t *)p; p += km_al((size_t)k * 4);
  double *partial = (double *)p; p += km_al((size_t)max_blocks * dim * 8);
  double *pw = (double *)p;
  ProfScope prof("km_update", st);
  int bits = 1;
  while ((1ll << bits) < k) ++bits;
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vals_in, n);
  cudaMemsetAsync(counts, 0, (size_t)k * 4, st);
  hist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(labels, n, counts);
  cub::DeviceRadixSort::SortPairs(d_tmp, sort_tmp, labels, keys_out, vals_in, vals_out, (int)n, 0, bits, st);  // stable
  cub::DeviceScan::ExclusiveSum(d_tmp, scan_tmp, counts, offs, k, st);
  kmeans_nblk_kernel<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(counts, k, nb);
  cub::DeviceScan::ExclusiveSum(d_tmp, scan_tmp, nb, poffs, k, st);
  if (n > KM_BLOCK)   // otherwise no cluster can be large
    kmeans_partial_kernel<T><<<(unsigned)max_blocks, dim >= 256 ? 256 : ((dim + 31) / 32) * 32, 0, st>>>(
        x, weights, dim, vals_out, offs, counts, nb, poffs, k, partial, pw);
  const int64_t total = (int64_t)k * dim;
  kmeans_update_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, weights, dim, vals_out, offs, counts, nb, poffs, partial, pw,
                                                                           k, nan_empty, divide, cent, counts_out, wsum_out);
  note_launch(11);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_kmeans_update_f64(const double *x, const double *weights, int64_t n, int dim, const int32_t *labels, int k, double *cent,
                             int64_t *counts_out, double *wsum_out, void *ws, size_t ws_bytes, int nan_empty, int divide,
                             cudaStream_t st) {
  return launch_kmeans_update_t<double>(x, weights, n, dim, labels, k, cent, counts_out, wsum_out, ws, ws_bytes, nan_empty, divide, st);
}
int launch_kmeans_update_i16(const int16_t *x, int64_t n, const int32_t *labels, int k, double *cent, int64_t *counts_out, void *ws,
                             size_t ws_bytes, int nan_empty, int divide, cudaStream_t st) {
  return launch_kmeans_update_t<int16_t>(x, nullptr, n, 192, labels, k, cent, counts_out, nullptr, ws, ws_bytes, nan_empty, divide, st);
}

// ------------------------------------------------------------------ tensor-core assignment for int16 points
// The candidate search runs on the exact int8-limb k-NN kernel against the centroids ROUNDED to int16; the exact f64
// distances to the KC nearest rounded centroids then decide the label.  With c = r + delta, |delta_i| <= 1/2:
//   |d_true - d_round| <= 2 sqrt(d_round) sqrt(192/4) + 192/4 = sqrt(192 d_round) + 48 =: eps(d_round),
// and d - eps(d) is increasing, so every centroid outside the candidates has d_true >= g(d_round of the KC-th candidate).
// If that exceeds the best exact candidate distance the label is certified; otherwise the point is queued for the
// brute-force f64 kernel.  Labels are therefore exactly the f64 arg-min (first minimum), like the oracle's.
__global__ void round_centroids_kernel(const double *__restrict__ cent, int64_t total, int16_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double c = cent[i];
  // empty clusters (NaN) are parked far away; they can never be the exact arg-min either (NaN compares false)
  out[i] = isnan(c) ? (int16_t)32767 : (int16_t)max(-32768, min(32767, __double2int_rn(c)));
}

template <int KC>
__global__ void __launch_bounds__(256)
kmeans_rerank_i16_kernel(const int16_t *__restrict__ x, int64_t n, const int32_t *__restrict__ cand, const uint32_t *__restrict__ cdist,
                         const double *__restrict__ cent, int k, int32_t *__restrict__ labels, double *__restrict__ dist,
                         int32_t *__restrict__ changed, int32_t *__restrict__ amb_list, int32_t *__restrict__ amb_count) {
  // KC threads per point (one candidate each): sequential 192-term f64 sums in the oracle's order
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = t / KC;
  const int c = (int)(t % KC);
  const bool live = p < n;
  const int32_t ci = live ? cand[p * KC + c] : -1;
  double s = INFINITY;
  if (ci >= 0 && ci < k) {
    const int16_t *xv = x + p * 192;
    const double *cv = cent + (int64_t)ci * 192;
    s = 0.0;
    for (int j = 0; j < 192; ++j) {
      const double df = __dsub_rn((double)__ldg(xv + j), __ldg(cv + j));
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    if (isnan(s)) s = INFINITY;
  }
  // arg-min over the KC lanes of this point: (distance, centroid index) lexicographic = first minimum in index order
  double bs = s;
  int32_t bi = (s < INFINITY) ? ci : 0x7fffffff;
#pragma unroll
  for (int o = KC / 2; o >= 1; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
  }
  if (live && c == 0) {
    // certificate from the largest rounded distance among the candidates (the list is sorted ascending)
    const int32_t last = cand[p * KC + KC - 1];
    bool certified;
    if (last < 0) certified = true;   // fewer than KC centroids exist: the candidates are all of them
    else {
      const double dl = (double)cdist[p * KC + KC - 1];
      certified = (dl - sqrt(192.0 * dl) - 48.0) > bs;
    }
    if (bi == 0x7fffffff) certified = false;
    if (certified) {
      if (labels[p] != bi) { labels[p] = bi; atomicAdd(changed, 1); }
      if (dist) dist[p] = bs;
    } else {
      amb_list[atomicAdd(amb_count, 1)] = (int32_t)p;
    }
  }
}

// second level for the uncertified points: their limb rows are gathered, the tensor-core search is repeated with 64
// candidates, and a warp per point evaluates the 64 exact f64 distances; the same bound with the 64th rounded distance
// certifies almost all of them.  Only the residue goes to the brute-force f64 kernel.
__global__ void amb_gather_limbs_kernel(const uint8_t *__restrict__ limbs, const uint32_t *__restrict__ norms,
                                        const int32_t *__restrict__ amb_list, int n_amb, uint8_t *__restrict__ out_limbs,
                                        uint32_t *__restrict__ out_norms) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one uint4 (16 bytes) per thread, 24 per row
  if (t >= (int64_t)n_amb * 24) return;
  const int i = (int)(t / 24), c = (int)(t % 24);
  const int64_t p = amb_list[i];
  reinterpret_cast<uint4 *>(out_limbs)[t] = reinterpret_cast<const uint4 *>(limbs + p * 384)[c];
  if (c == 0) out_norms[i] = norms[p];
}

__global__ void __launch_bounds__(256)
kmeans_rerank64_kernel(const int16_t *__restrict__ x, const int32_t *__restrict__ amb_list, int n_amb, const int32_t *__restrict__ cand,
                       const uint32_t *__restrict__ cdist, const double *__restrict__ cent, int k, int32_t *__restrict__ labels,
                       double *__restrict__ dist, int32_t *__restrict__ changed, int32_t *__restrict__ amb2_list,
                       int32_t *__restrict__ amb2_count) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_amb) return;
  const int64_t p = amb_list[i];
  const int16_t *xv = x + p * 192;
  double bs = INFINITY;
  int32_t bi = 0x7fffffff;
  for (int c = lane; c < 64; c += 32) {
    const int32_t ci = cand[(int64_t)i * 64 + c];
    if (ci < 0 || ci >= k) continue;
    const double *cv = cent + (int64_t)ci * 192;
    double s = 0.0;
    for (int j = 0; j < 192; ++j) {
      const double df = __dsub_rn((double)__ldg(xv + j), __ldg(cv + j));
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    if (s < bs || (s == bs && ci < bi)) { bs = s; bi = ci; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
  }
  if (lane == 0) {
    const int32_t last = cand[(int64_t)i * 64 + 63];
    bool certified;
    if (last < 0) certified = true;   // fewer than 64 centroids: the candidates are all of them
    else {
      const double dl = (double)cdist[(int64_t)i * 64 + 63];
      certified = (dl - sqrt(192.0 * dl) - 48.0) > bs;
    }
    if (bi == 0x7fffffff) certified = false;
    if (certified) {
      if (labels[p] != bi) { labels[p] = bi; atomicAdd(changed, 1); }
      if (dist) dist[p] = bs;
    } else {
      amb2_list[atomicAdd(amb2_count, 1)] = (int32_t)p;
    }
  }
}

// queued (uncertified) points: gathered into a dense f64 matrix, assigned by the exact f64 kernel (32 points per block
// share every centroid load), scattered back
__global__ void amb_gather_kernel(const int16_t *__restrict__ x, const int32_t *__restrict__ amb_list, int n_amb,
                                  const int32_t *__restrict__ labels, double *__restrict__ xa, int32_t *__restrict__ la) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_amb * 192) return;
  const int i = (int)(t / 192), j = (int)(t % 192);
  const int64_t p = amb_list[i];
  xa[t] = (double)x[p * 192 + j];
  if (j == 0) la[i] = labels[p];
}
__global__ void amb_scatter_kernel(const int32_t *__restrict__ amb_list, int n_amb, const int32_t *__restrict__ la,
                                   const double *__restrict__ da, int32_t *__restrict__ labels, double *__restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_amb) return;
  const int64_t p = amb_list[i];
  labels[p] = la[i];
  if (dist) dist[p] = da[i];
}

int launch_amb_gather_limbs(const uint8_t *limbs, const uint32_t *norms, const int32_t *amb_list, int n_amb, uint8_t *out_limbs,
                            uint32_t *out_norms, cudaStream_t st) {
  if (n_amb <= 0) return TM_OK;
  const int64_t total = (int64_t)n_amb * 24;
  amb_gather_limbs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(limbs, norms, amb_list, n_amb, out_limbs, out_norms);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}
int launch_kmeans_rerank64(const int16_t *x, const int32_t *amb_list, int n_amb, const int32_t *cand, const uint32_t *cdist,
                           const double *cent, int k, int32_t *labels, double *dist, int32_t *changed, int32_t *amb2_list,
                           int32_t *amb2_count, cudaStream_t st) {
  if (n_amb <= 0) return TM_OK;
  ProfScope prof("km_rerank64", st);
  kmeans_rerank64_kernel<<<(n_amb + 7) / 8, 256, 0, st>>>(x, amb_list, n_amb, cand, cdist, cent, k, labels, dist, changed, amb2_list, amb2_count);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_round_centroids(const double *cent, int k, int16_t *out, cudaStream_t st) {
  const int64_t total = (int64_t)k * 192;
  round_centroids_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cent, total, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_kmeans_rerank_i16(const int16_t *x, int64_t n, const int32_t *cand, const uint32_t *cdist, const double *cent, int k,
                             int32_t *labels, double *dist, int32_t *changed, int32_t *amb_list, int32_t *amb_count, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  ProfScope prof("km_rerank", st);
  const int64_t threads = n * KMEANS_KC;
  kmeans_rerank_i16_kernel<KMEANS_KC><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, n, cand, cdist, cent, k, labels, dist, changed,
                                                                                        amb_list, amb_count);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_kmeans_assign_amb(const int16_t *x, const int32_t *amb_list, int n_amb, const double *cent, int k, int32_t *labels,
                             double *dist, int32_t *changed, double *xa, int32_t *la, double *da, cudaStream_t st) {
  if (n_amb <= 0) return TM_OK;
  ProfScope prof("km_amb", st);
  const int64_t total = (int64_t)n_amb * 192;
  amb_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, amb_list, n_amb, labels, xa, la);
  note_launch();
  int rc = launch_kmeans_assign_f64(xa, n_amb, 192, cent, k, la, da, changed, st);
  if (rc) return rc;
  amb_scatter_kernel<<<(n_amb + 255) / 256, 256, 0, st>>>(amb_list, n_amb, la, da, labels, dist);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_kmeanspp_f64(const double *x, int64_t n, int dim, int k, unsigned long long seed, double *d2_ws, double *cent,
                        cudaStream_t st) {
  if (n <= 0 || k < 1 || dim < 1 || dim > 4096) return TM_ERR_ARG;
  unsigned long long *state = nullptr;
  if (cudaMallocAsync(&state, 8, st) != cudaSuccess) return TM_ERR_NOMEM;
  kmpp_pick_kernel<<<1, 256, 0, st>>>(x, n, dim, 0, seed, d2_ws, state, cent);
  for (int c = 1; c < k; ++c) {
    kmpp_dist_kernel<<<(unsigned)((n + 127) / 128), 128, (size_t)dim * sizeof(double), st>>>(x, n, dim, cent + (int64_t)(c - 1) * dim, c == 1, d2_ws);
    kmpp_pick_kernel<<<1, 256, 0, st>>>(x, n, dim, c, seed, d2_ws, state, cent);
  }
  cudaFreeAsync(state, st);
  note_launch(2 * k - 1);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_kmeans_finish(const double *sums, const int64_t *counts, int k, int dim, int nan_empty, double *cent, cudaStream_t st) {
  const int64_t total = (int64_t)k * dim;
  kmeans_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sums, counts, k, dim, nan_empty, cent);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// ------------------------------------------------------------------ RGB (3-d) k-means, all palettes at once
// pixels: packed 0x00BBGGRR grouped by palette (segment p = [off[p], off[p+1])), cent [n_pal][k][3] f64.
// One Lloyd iteration = assign+accumulate kernel, then a centroid kernel.
struct RgbAcc { unsigned long long r, g, b, n; };

__global__ void __launch_bounds__(256)
rgb_assign_kernel(const int32_t *__restrict__ px, const int64_t *__restrict__ off, int n_pal, int k, const double *__restrict__ cent,
                  uint16_t *__restrict__ labels, RgbAcc *__restrict__ acc, int32_t *__restrict__ changed, int64_t chunk,
                  const int32_t *__restrict__ kcount) {
  extern __shared__ uint8_t s_mem[];
  double *s_c = reinterpret_cast<double *>(s_mem);                     // [k][3]
  unsigned int *s_a = reinterpret_cast<unsigned int *>(s_c + 3 * k);  // [k][4]
  const int64_t n_total = off[n_pal];
  const int64_t c0 = (int64_t)blockIdx.x * chunk, c1 = min(c0 + chunk, n_total);
  if (c0 >= c1) return;
  // first palette whose segment ends after c0
  int p = 0;
  {
    int lo = 0, hi = n_pal - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (off[mid + 1] > c0) hi = mid; else lo = mid + 1; }
    p = lo;
  }
  int local_changed = 0;
  for (; p < n_pal && off[p] < c1; ++p) {
    const int64_t s0 = max(c0, off[p]), s1 = min(c1, off[p + 1]);
    if (s0 >= s1) continue;
    const int kk = kcount[p];  // slots beyond Min(AColorCount, DSLen) do not exist (:4464)
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * k; i += 256) s_c[i] = cent[(int64_t)p * k * 3 + i];
    for (int i = threadIdx.x; i < 4 * k; i += 256) s_a[i] = 0;
    __syncthreads();
    for (int64_t i = s0 + threadIdx.x; i < s1; i += 256) {
      const int32_t c = __ldg(px + i);
      const double r = (double)(c & 255), g = (double)((c >> 8) & 255), b = (double)((c >> 16) & 255);
      double best = INFINITY;
      int bi = -1;
      for (int j = 0; j < kk; ++j) {
        const double dr = __dsub_rn(r, s_c[3 * j]), dg = __dsub_rn(g, s_c[3 * j + 1]), db = __dsub_rn(b, s_c[3 * j + 2]);
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dg, dg)), __dmul_rn(db, db));
        if (s < best) { best = s; bi = j; }
      }
      const int old = labels[i];
      if (bi < 0) bi = old != 0xFFFF ? old : 0;
      if (bi != old) { labels[i] = (uint16_t)bi; ++local_changed; }
      atomicAdd(&s_a[4 * bi], (unsigned)(c & 255));
      atomicAdd(&s_a[4 * bi + 1], (unsigned)((c >> 8) & 255));
      atomicAdd(&s_a[4 * bi + 2], (unsigned)((c >> 16) & 255));
      atomicAdd(&s_a[4 * bi + 3], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += 256) {
      if (s_a[4 * j + 3]) {
        RgbAcc *a = acc + (int64_t)p * k + j;
        atomicAdd(&a->r, (unsigned long long)s_a[4 * j]);
        atomicAdd(&a->g, (unsigned long long)s_a[4 * j + 1]);
        atomicAdd(&a->b, (unsigned long long)s_a[4 * j + 2]);
        atomicAdd(&a->n, (unsigned long long)s_a[4 * j + 3]);
      }
    }
  }
  if (local_changed) atomicAdd(changed, local_changed);
}

__global__ void rgb_centroid_kernel(const RgbAcc *__restrict__ acc, int64_t total, double *__restrict__ cent, const int32_t *__restrict__ kcount,
                                    int k) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i / k), j = (int)(i % k);
  if (j >= kcount[p]) return;  // unused slots (k clipped to the pixel count)
  const RgbAcc a = acc[i];
  if (a.n) {
    cent[3 * i] = __ddiv_rn((double)a.r, (double)a.n);
    cent[3 * i + 1] = __ddiv_rn((double)a.g, (double)a.n);
    cent[3 * i + 2] = __ddiv_rn((double)a.b, (double)a.n);
  } else {
    cent[3 * i] = NAN; cent[3 * i + 1] = NAN; cent[3 * i + 2] = NAN;  // empty cluster (host tolerates NaN, :4521)
  }
}

// k-means++ seeding identical to the oracle's tmo_kmeanspp_init (explicit xorshift64* stream, D^2 sampling by first
// prefix sum exceeding u).  Integer distances -> prefix sums are exact in any order.  One 1024-thread block per palette:
// warp w owns a contiguous range of the palette's pixels and reads it 32 consecutive pixels at a time (coalesced; the first
// version gave every THREAD a contiguous range -- 32 cache lines per warp load, 22 ms for 16 palettes x 262 144 pixels); the
// pick is located by the owner warp with a warp-wide inclusive scan over its range, 32 pixels per step.
constexpr int KPP_THREADS = 1024, KPP_WARPS = KPP_THREADS / 32;
__global__ void __launch_bounds__(KPP_THREADS)
rgb_kmeanspp_kernel(const int32_t *__restrict__ px, const int64_t *__restrict__ off, int k, const int32_t *__restrict__ kcount,
                    unsigned long long seed, unsigned int *__restrict__ d2, double *__restrict__ cent) {
  const int p = blockIdx.x;
  const int64_t s0 = off[p], n = off[p + 1] - s0;
  const int kk = kcount[p];
  if (n <= 0 || kk <= 0) return;
  __shared__ unsigned long long s_part[KPP_WARPS];
  __shared__ unsigned long long s_acc;     // pixels before the owner warp's range: their summed distances
  __shared__ double s_u;
  __shared__ int s_owner;
  __shared__ int s_last[3];
  unsigned long long st = seed ? seed : 0x9E3779B97F4A7C15ULL;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int64_t per = (n + KPP_WARPS - 1) / KPP_WARPS, a = min((int64_t)w * per, n), b = min(a + per, n);
  if (t == 0) {
    const int64_t first = (int64_t)(xorshift64s(st) % (unsigned long long)n);
    const int32_t c = px[s0 + first];
    s_last[0] = c & 255; s_last[1] = (c >> 8) & 255; s_last[2] = (c >> 16) & 255;
    double *cv = cent + ((int64_t)p * k) * 3;
    cv[0] = s_last[0]; cv[1] = s_last[1]; cv[2] = s_last[2];
  }
  for (int64_t i = a + lane; i < b; i += 32) d2[s0 + i] = 0xFFFFFFFFu;
  for (int c = 1; c < kk; ++c) {
    __syncthreads();
    const int lr = s_last[0], lg = s_last[1], lb = s_last[2];
    unsigned long long part = 0;
    for (int64_t i = a + lane; i < b; i += 32) {
      const int32_t v = px[s0 + i];
      const int dr = (v & 255) - lr, dg = ((v >> 8) & 255) - lg, db = ((v >> 16) & 255) - lb;
      const unsigned int s = (unsigned)(dr * dr + dg * dg + db * db);
      unsigned int cur = d2[s0 + i];
      if (s < cur) { cur = s; d2[s0 + i] = s; }
      part += cur;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_part[w] = part;
    __syncthreads();
    if (t == 0) {
      unsigned long long total = 0;
      for (int i = 0; i < KPP_WARPS; ++i) total += s_part[i];
      const double u = (double)(xorshift64s(st) >> 11) * (1.0 / 9007199254740992.0) * (double)total;
      // owner warp = first whose inclusive prefix exceeds u (none: the pick falls back to the last pixel, like the oracle's scan)
      unsigned long long acc = 0;
      int owner = -1;
      for (int i = 0; i < KPP_WARPS; ++i) {
        if ((double)(acc + s_part[i]) > u) { owner = i; break; }
        acc += s_part[i];
      }
      s_owner = owner; s_acc = acc; s_u = u;
      if (owner < 0) {
        const int32_t v = px[s0 + n - 1];
        s_last[0] = v & 255; s_last[1] = (v >> 8) & 255; s_last[2] = (v >> 16) & 255;
        double *cv = cent + ((int64_t)p * k + c) * 3;
        cv[0] = s_last[0]; cv[1] = s_last[1]; cv[2] = s_last[2];
      }
    }
    __syncthreads();
    if (w == s_owner) {   // first pixel of this warp's range whose inclusive prefix exceeds u
      unsigned long long acc = s_acc;
      const double u = s_u;
      int64_t pick = n - 1;
      for (int64_t j = a; j < b; j += 32) {
        const int64_t i = j + lane;
        unsigned long long incl = i < b ? (unsigned long long)d2[s0 + i] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, i < b && (double)(acc + incl) > u);
        if (hit) { pick = j + (__ffs(hit) - 1); break; }
        acc += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) {
        const int32_t v = px[s0 + pick];
        s_last[0] = v & 255; s_last[1] = (v >> 8) & 255; s_last[2] = (v >> 16) & 255;
        double *cv = cent + ((int64_t)p * k + c) * 3;
        cv[0] = s_last[0]; cv[1] = s_last[1]; cv[2] = s_last[2];
      }
    }
  }
}

// centroids -> palette (round half-even, clamp, integer HSV, sort by (V,S,H), pad with the null colour)
__device__ int muldiv_rn(int a, int b, int c) {  // Win32 MulDiv
  if (c == 0) return -1;
  const long long p = (long long)a * b;
  const long long ac = c < 0 ? -(long long)c : c, ap = p < 0 ? -p : p;
  const long long q = (ap + ac / 2) / ac;
  return (int)(((p < 0) != (c < 0)) ? -q : q);
}
__device__ void rgb_to_hsv(int rr, int gg, int bb, int &h, int &s, int &v) {  // utils.pas:278-325
  const int mx = max(rr, max(gg, bb)), mn = min(rr, min(gg, bb));
  int hh = 0, ss = 0;
  if (mx != mn) {
    const int delta = mx - mn;
    ss = muldiv_rn(delta, 255, mx);
    if (rr == mx) hh = muldiv_rn(42, gg - bb, delta);
    else if (gg == mx) hh = muldiv_rn(42, bb - rr, delta) + 84;
    else hh = muldiv_rn(42, rr - gg, delta) + 168;
    hh = hh % 252;
  }
  h = hh & 255; s = ss & 255; v = mx & 255;
}

__global__ void palette_finish_kernel(const double *__restrict__ cent, const int32_t *__restrict__ kcount, int n_pal, int k,
                                      int pal_size, int32_t *__restrict__ palettes) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pal) return;
  int32_t *out = palettes + (int64_t)p * pal_size;
  const int kk = kcount[p];
  // insertion sort on keys (V,S,H,order) kept in the output row itself plus a key array in local memory
  unsigned int keys[256];
  for (int i = 0; i < kk; ++i) {
    const double *c = cent + ((int64_t)p * k + i) * 3;
    int r = 0, g = 0, b = 0;
    if (!isnan(c[0]) && !isnan(c[1]) && !isnan(c[2])) {
      r = min(max(__double2int_rn(c[0]), 0), 255);
      g = min(max(__double2int_rn(c[1]), 0), 255);
      b = min(max(__double2int_rn(c[2]), 0), 255);
    }
    int h, s, v;
    rgb_to_hsv(r, g, b, h, s, v);
    const unsigned int key = ((unsigned)v << 24) | ((unsigned)s << 16) | ((unsigned)h << 8) | (unsigned)i;
    const int32_t col = (b << 16) | (g << 8) | r;
    int j = i;
    while (j > 0 && keys[j - 1] > key) { keys[j] = keys[j - 1]; out[j] = out[j - 1]; --j; }
    keys[j] = key;
    out[j] = col;
  }
  for (int i = kk; i < pal_size; ++i) out[i] = (int32_t)0xffff00ff;
}

// keys for grouping pixels by palette and ordering them (G,R,B) like CompareDSPixel (tilingencoder.pas:1046-1056)
__global__ void pixel_keys_kernel(const int32_t *__restrict__ rgb, const int32_t *__restrict__ tile_pal, int64_t n_tiles,
                                  unsigned long long *__restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles * 64) return;
  const int32_t c = rgb[i];
  const unsigned int r = c & 255, g = (c >> 8) & 255, b = (c >> 16) & 255;
  keys[i] = ((unsigned long long)(unsigned)tile_pal[i >> 6] << 24) | (g << 16) | (r << 8) | b;
}
__global__ void keys_to_pixels_kernel(const unsigned long long *__restrict__ keys, int64_t n, int32_t *__restrict__ px,
                                      unsigned long long *__restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long kx = keys[i];
  const unsigned int g = (kx >> 16) & 255, r = (kx >> 8) & 255, b = kx & 255;
  px[i] = (int32_t)((b << 16) | (g << 8) | r);
  atomicAdd(counts + (kx >> 24), 1ull);
}

// Full colour-quantisation stage for all palettes (QuantizeUsingYakmo + DoQuantization over PreparePalettes' loop,
// tilingencoder.pas:1864, 4434-4564).  init: [n_pal][pal_size][3] explicit centroids or nullptr (seeded k-means++).
int run_palette_quantise(const int32_t *rgb, const int32_t *tile_pal, int64_t n_tiles, int n_pal, int pal_size, const double *init,
                         unsigned long long seed, int max_iter, int32_t *palettes_out, int32_t *iters_out, cudaStream_t st) {
  if (n_tiles <= 0 || n_pal < 1 || pal_size < 1 || pal_size > 256) return TM_ERR_ARG;
  const int64_t n = n_tiles * 64;
  const int k = pal_size;
  int bits = 24;
  while ((1ll << (bits - 24)) < n_pal) ++bits;
  unsigned long long *keys = nullptr, *keys2 = nullptr, *counts = nullptr;
  int32_t *px = nullptr, *kcount = nullptr, *changed = nullptr;
  int64_t *off = nullptr;
  double *cent = nullptr;
  uint16_t *labels = nullptr;
  RgbAcc *acc = nullptr;
  unsigned int *d2 = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = TM_OK;
  std::vector<unsigned long long> h_counts(n_pal);
  std::vector<int64_t> h_off(n_pal + 1);
  std::vector<int32_t> h_kcount(n_pal);
  int it = 0;
#define CK(x) do { if ((x) != cudaSuccess) { rc = TM_ERR_CUDA; goto done; } } while (0)
  // stream-ordered scratch: the pool keeps it between calls (plain cudaMalloc / cudaFree cost 0.6 s per call here)
  CK(cudaMallocAsync(&keys, n * 8, st)); CK(cudaMallocAsync(&keys2, n * 8, st)); CK(cudaMallocAsync(&counts, (size_t)n_pal * 8, st));
  CK(cudaMallocAsync(&px, n * 4, st)); CK(cudaMallocAsync(&kcount, (size_t)n_pal * 4, st)); CK(cudaMallocAsync(&changed, 4, st));
  CK(cudaMallocAsync(&off, (size_t)(n_pal + 1) * 8, st)); CK(cudaMallocAsync(&cent, (size_t)n_pal * k * 3 * 8, st));
  CK(cudaMallocAsync(&labels, n * 2, st)); CK(cudaMallocAsync(&acc, (size_t)n_pal * k * sizeof(RgbAcc), st)); CK(cudaMallocAsync(&d2, n * 4, st));
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, keys2, (int)n, 0, bits, st);
  CK(cudaMallocAsync(&tmp, tmp_bytes, st));
  note_launch(6);
  pixel_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rgb, tile_pal, n_tiles, keys);
  cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys, keys2, (int)n, 0, bits, st);
  CK(cudaMemsetAsync(counts, 0, (size_t)n_pal * 8, st));
  keys_to_pixels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys2, n, px, counts);
  CK(cudaMemcpyAsync(h_counts.data(), counts, (size_t)n_pal * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h_off[0] = 0;
  for (int p = 0; p < n_pal; ++p) {
    h_off[p + 1] = h_off[p] + (int64_t)h_counts[p];
    h_kcount[p] = (int32_t)((int64_t)h_counts[p] < k ? (int64_t)h_counts[p] : k);  // AColorCount := Min(AColorCount, DSLen), :4464
  }
  CK(cudaMemcpyAsync(off, h_off.data(), (size_t)(n_pal + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(kcount, h_kcount.data(), (size_t)n_pal * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(labels, 0xFF, n * 2, st));
  if (init) CK(cudaMemcpyAsync(cent, init, (size_t)n_pal * k * 3 * 8, cudaMemcpyDefault, st));
  else {
    CK(cudaMemsetAsync(cent, 0xFF, (size_t)n_pal * k * 3 * 8, st));  // NaN fill: unused slots never win
    rgb_kmeanspp_kernel<<<n_pal, KPP_THREADS, 0, st>>>(px, off, k, kcount, seed, d2, cent);
  }
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t chunk = (n + (int64_t)sms * 8 - 1) / ((int64_t)sms * 8);
    if (chunk < 2048) chunk = 2048;
    if (chunk > 16 * 1024 * 1024) chunk = 16 * 1024 * 1024;  // 255 * chunk must fit the 32-bit block accumulators
    const unsigned grid = (unsigned)((n + chunk - 1) / chunk);
    const size_t smem = (size_t)k * (3 * 8 + 4 * 4);
    // Lloyd: a palette with k == 1 is a plain mean (:4502-4509) = one update from any start
    for (;;) {
      int32_t h_changed = 0;
      CK(cudaMemsetAsync(changed, 0, 4, st));
      CK(cudaMemsetAsync(acc, 0, (size_t)n_pal * k * sizeof(RgbAcc), st));
      note_launch(2);
      rgb_assign_kernel<<<grid, 256, smem, st>>>(px, off, n_pal, k, cent, labels, acc, changed, chunk, kcount);
      CK(cudaMemcpyAsync(&h_changed, changed, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (h_changed == 0 || it >= max_iter) break;
      ++it;
      rgb_centroid_kernel<<<(unsigned)(((int64_t)n_pal * k + 255) / 256), 256, 0, st>>>(acc, (int64_t)n_pal * k, cent, kcount, k);
    }
  }
  palette_finish_kernel<<<(n_pal + 63) / 64, 64, 0, st>>>(cent, kcount, n_pal, k, pal_size, palettes_out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  if (iters_out) *iters_out = it;
done:
#undef CK
  for (void *ptr : {(void *)keys, (void *)keys2, (void *)counts, (void *)px, (void *)kcount, (void *)changed, (void *)off, (void *)cent,
                    (void *)labels, (void *)acc, (void *)d2, tmp})
    if (ptr) cudaFreeAsync(ptr, st);
  return rc;
}

}  // namespace tmg
