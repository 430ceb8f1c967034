// reduce.cu -- exact duplicate detection over millions of 256-byte tiles: the MakeTilesUnique(True) step inside the
// tile-count search of TTilingEncoder.Reduce (STCGREval / TransferTiles / MakeTilesUnique, tilingencoder.pas:4014-4103,
// 4720-4781).  HBM-bound: every tile is read twice (hash, verification).
//
//   1. tile_hash_kernel      : warp per tile, two independent 64-bit position-salted hashes.
//   2. two stable radix sorts: by h2, then by h1  ->  tiles ordered by the 128-bit key.
//   3. class_boundary_kernel : a sorted tile opens a new class when its key OR its 64 pixels differ from its predecessor's
//                              (so equal tiles always share a class; the pixel compare makes a 128-bit collision harmless
//                              unless three tiles interleave, probability ~1e-25).
//   4. inclusive scan of the boundary flags -> class id, scattered back to the tiles' own order.
#include "tm_kernels.h"
#include <cub/cub.cuh>

namespace tmg {

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256) tile_hash_kernel(const int32_t *__restrict__ rgb, int64_t n, unsigned long long *__restrict__ h1,
                                                        unsigned long long *__restrict__ h2, int32_t *__restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tile >= n) return;
  const uint2 v = __ldg(reinterpret_cast<const uint2 *>(rgb + tile * 64) + lane);
  const unsigned long long w = ((unsigned long long)v.y << 32) | v.x;
  unsigned long long a = mix64(w + 0x9E3779B97F4A7C15ULL * (unsigned long long)(lane + 1));
  unsigned long long b = mix64(w * 0xA24BAED4963EE407ULL + (unsigned long long)lane * 0x9FB21C651E98DF25ULL + 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {   // sums of position-salted mixes: order independent across lanes, position dependent
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) { h1[tile] = a; h2[tile] = b; idx[tile] = (int32_t)tile; }
}

__global__ void __launch_bounds__(256) gather_u64_kernel(const unsigned long long *__restrict__ src, const int32_t *__restrict__ idx,
                                                         int64_t n, unsigned long long *__restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(256) class_boundary_kernel(const int32_t *__restrict__ rgb, const int32_t *__restrict__ sorted_idx,
                                                             const unsigned long long *__restrict__ h1s,
                                                             const unsigned long long *__restrict__ h2, int64_t n,
                                                             int32_t *__restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  bool differ = true;
  if (i > 0) {
    const int32_t a = sorted_idx[i], b = sorted_idx[i - 1];
    differ = h1s[i] != h1s[i - 1] || h2[a] != h2[b];
    if (!differ) {
      const uint2 x = __ldg(reinterpret_cast<const uint2 *>(rgb + (int64_t)a * 64) + lane);
      const uint2 y = __ldg(reinterpret_cast<const uint2 *>(rgb + (int64_t)b * 64) + lane);
      differ = __any_sync(0xffffffffu, x.x != y.x || x.y != y.y);
    }
  }
  if (lane == 0) flag[i] = differ ? 1 : 0;
}

__global__ void __launch_bounds__(256) scatter_class_kernel(const int32_t *__restrict__ sorted_idx, const int32_t *__restrict__ scan,
                                                            int64_t n, int32_t *__restrict__ class_id) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) class_id[sorted_idx[i]] = scan[i] - 1;
}

size_t tile_classes_ws_bytes(int64_t n) {
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  const size_t tmp = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  return ((tmp + 255) & ~(size_t)255) + (size_t)n * (8 * 4 + 4 * 4) + 1024;
}

// class_id[n] (0 .. n_classes-1, numbered in hash order); *n_classes_dev receives the class count (device int32)
int run_tile_classes(const int32_t *rgb, int64_t n, int32_t *class_id, int32_t *n_classes_dev, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0 || n > 0x7fffffff) return TM_ERR_ARG;
  if (ws_bytes < tile_classes_ws_bytes(n)) return TM_ERR_ARG;
  size_t sort_tmp = 0, scan_tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
  size_t tmp_bytes = sort_tmp > scan_tmp ? sort_tmp : scan_tmp;
  uint8_t *p = (uint8_t *)ws;
  void *d_tmp = p; p += (tmp_bytes + 255) & ~(size_t)255;
  unsigned long long *h1 = (unsigned long long *)p; p += (size_t)n * 8;
  unsigned long long *h2 = (unsigned long long *)p; p += (size_t)n * 8;
  unsigned long long *ka = (unsigned long long *)p; p += (size_t)n * 8;
  unsigned long long *kb = (unsigned long long *)p; p += (size_t)n * 8;
  int32_t *ia = (int32_t *)p; p += (size_t)n * 4;
  int32_t *ib = (int32_t *)p; p += (size_t)n * 4;
  int32_t *flag = (int32_t *)p; p += (size_t)n * 4;
  int32_t *scan = (int32_t *)p;
  ProfScope prof("tile_classes", st);
  const unsigned wb = (unsigned)((n + 7) / 8), tb = (unsigned)((n + 255) / 256);
  tile_hash_kernel<<<wb, 256, 0, st>>>(rgb, n, h1, h2, ia);
  // stable LSD over the 128-bit key: sort by h2 first, then by h1
  if (cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, h2, ka, ia, ib, (int)n, 0, 64, st) != cudaSuccess) return TM_ERR_CUDA;
  gather_u64_kernel<<<tb, 256, 0, st>>>(h1, ib, n, kb);
  if (cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kb, ka, ib, ia, (int)n, 0, 64, st) != cudaSuccess) return TM_ERR_CUDA;
  // ia = tile indices in key order, ka = their h1
  class_boundary_kernel<<<wb, 256, 0, st>>>(rgb, ia, ka, h2, n, flag);
  if (cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, flag, scan, (int)n, st) != cudaSuccess) return TM_ERR_CUDA;
  scatter_class_kernel<<<tb, 256, 0, st>>>(ia, scan, n, class_id);
  if (cudaMemcpyAsync(n_classes_dev, scan + (n - 1), 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) return TM_ERR_CUDA;
  note_launch(7);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// ------------------------------------------------------------------ tile-count search bookkeeping (STCGREval, :4014-4046)
// Per duplicate class the SMALLEST effective PSNR of its members decides from which threshold x on the class contributes a
// dictionary tile (a tile is unpredicted when not PSNR > x), so the count of distinct unpredicted tiles for any x is a binary
// search in the sorted class minima: the golden-ratio search (utils.pas:1044-1072) then needs no further pass over the tiles.
// All three kernels are one pass over n tiles with atomics on per-class slots: HBM-bound (12-16 bytes per tile).
__device__ __forceinline__ unsigned long long f64_order_key(double v) {   // monotone map double -> uint64 (negative values included)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double f64_from_order_key(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long *__restrict__ p, int64_t n, unsigned long long v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void __launch_bounds__(256) class_min_kernel(const int32_t *__restrict__ cls, const double *__restrict__ eff, int64_t n,
                                                        unsigned long long *__restrict__ key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMin(key + cls[i], f64_order_key(eff[i]));
}
__global__ void __launch_bounds__(256) keys_to_f64_kernel(const unsigned long long *__restrict__ key, int64_t n, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f64_from_order_key(key[i]);
}
size_t class_min_ws_bytes(int64_t n_cls) {
  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n_cls);
  return ((sort_tmp + 255) & ~(size_t)255) + (size_t)n_cls * 16 + 256;
}
// sorted_min[n_cls]: the classes' minimum effective PSNR, ascending
int run_class_min_sorted(const int32_t *cls, const double *eff, int64_t n, int64_t n_cls, double *sorted_min, void *ws, size_t ws_bytes,
                         cudaStream_t st) {
  if (n <= 0 || n_cls <= 0 || n_cls > 0x7fffffff || ws_bytes < class_min_ws_bytes(n_cls)) return TM_ERR_ARG;
  size_t sort_tmp = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_tmp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)n_cls);
  uint8_t *p = (uint8_t *)ws;
  void *d_tmp = p; p += (sort_tmp + 255) & ~(size_t)255;
  unsigned long long *ka = (unsigned long long *)p; p += (size_t)n_cls * 8;
  unsigned long long *kb = (unsigned long long *)p;
  ProfScope prof("reduce_class_min", st);
  fill_u64_kernel<<<(unsigned)((n_cls + 255) / 256), 256, 0, st>>>(ka, n_cls, ~0ull);
  class_min_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cls, eff, n, ka);
  if (cub::DeviceRadixSort::SortKeys(d_tmp, sort_tmp, ka, kb, (int)n_cls, 0, 64, st) != cudaSuccess) return TM_ERR_CUDA;
  keys_to_f64_kernel<<<(unsigned)((n_cls + 255) / 256), 256, 0, st>>>(kb, n_cls, sorted_min);
  note_launch(4);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// IsPredicted := PSNR > x (:4028-4031); per class: number of unpredicted members (the dictionary tile's UseCount after
// MakeTilesUnique, :4783-4815) and the first unpredicted member in frame order (the representative that is transferred)
__global__ void __launch_bounds__(256) reduce_apply_kernel(const int32_t *__restrict__ cls, const double *__restrict__ eff, int64_t n, double x,
                                                           int32_t *__restrict__ use, int32_t *__restrict__ rep, uint8_t *__restrict__ unpred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool u = !(eff[i] > x);
  unpred[i] = u ? 1 : 0;
  if (u) {
    const int32_t c = cls[i];
    atomicAdd(use + c, 1);
    atomicMin(rep + c, (int32_t)i);
  }
}
__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t *__restrict__ p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int run_reduce_apply(const int32_t *cls, const double *eff, int64_t n, int64_t n_cls, double x, int32_t *use, int32_t *rep, uint8_t *unpred,
                     cudaStream_t st) {
  if (n <= 0 || n > 0x7fffffff || n_cls <= 0) return TM_ERR_ARG;
  ProfScope prof("reduce_apply", st);
  fill_i32_kernel<<<(unsigned)((n_cls + 255) / 256), 256, 0, st>>>(use, n_cls, 0);
  fill_i32_kernel<<<(unsigned)((n_cls + 255) / 256), 256, 0, st>>>(rep, n_cls, (int32_t)n);
  reduce_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cls, eff, n, x, use, rep, unpred);
  note_launch(3);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// TMI.TileIdx := new index of the tile's class when unpredicted, -1 otherwise (TransferTiles + ReindexTiles' remap, :4066-4090, :4682-4693)
__global__ void __launch_bounds__(256) reduce_remap_kernel(const int32_t *__restrict__ cls, const uint8_t *__restrict__ unpred,
                                                           const int32_t *__restrict__ new_of_cls, int64_t n, int32_t *__restrict__ tile_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tile_idx[i] = unpred[i] ? new_of_cls[cls[i]] : -1;
}
int run_reduce_remap(const int32_t *cls, const uint8_t *unpred, const int32_t *new_of_cls, int64_t n, int32_t *tile_idx, cudaStream_t st) {
  if (n <= 0) return TM_ERR_ARG;
  reduce_remap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cls, unpred, new_of_cls, n, tile_idx);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
