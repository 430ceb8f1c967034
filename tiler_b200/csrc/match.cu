// match.cu -- the matcher's exact integer pieces around the tensor-core k-NN:
//   * CompareEuclideanDCTPtr (utils.pas:541-557) for explicit vector pairs,
//   * the "extended palette usage" re-rank of TFrame.Reconstruct.DoXY (tilingencoder.pas:1563-1609): among the k
//     nearest tiles, every unique tile recoloured with every unique palette of those tiles, strict minimum in
//     ascending (tile, palette) order,
//   * exact nearest neighbour of f64 vectors (the ANN.dll contract, extern.pas:178-180; call site :4128).
#include "tm_kernels.h"

namespace tmg {

__device__ __forceinline__ uint32_t sqdiff2(uint32_t a, uint32_t b) {
  const int a0 = (int16_t)(a & 0xffff), a1 = (int16_t)(a >> 16);
  const int b0 = (int16_t)(b & 0xffff), b1 = (int16_t)(b >> 16);
  const int d0 = a0 - b0, d1 = a1 - b1;
  return (uint32_t)(d0 * d0) + (uint32_t)(d1 * d1);
}

// one warp per pair: out[i] = sum (a-b)^2 mod 2^32
__global__ void __launch_bounds__(256) distance_pairs_kernel(const int16_t *__restrict__ a, const int16_t *__restrict__ b, int64_t n,
                                                            uint32_t *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint32_t *pa = reinterpret_cast<const uint32_t *>(a + i * 192), *pb = reinterpret_cast<const uint32_t *>(b + i * 192);
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) s += sqdiff2(__ldg(pa + lane + 32 * j), __ldg(pb + lane + 32 * j));
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) out[i] = s;
}

// sum of squares mod 2^32 of int16 rows [n][192]: one warp per row (norms of the (tile x palette) feature table)
__global__ void __launch_bounds__(256) row_norms_kernel(const int16_t *__restrict__ a, int64_t n, uint32_t *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint32_t *pa = reinterpret_cast<const uint32_t *>(a + i * 192);
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) s += sqdiff2(__ldg(pa + lane + 32 * j), 0u);
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) out[i] = s;
}

// q . b for two packed int16 pairs through the byte dot-product unit: b = 256 bh + bl (bh signed, bl unsigned), bytes of the
// permuted word p = [bl0, bl1, bh0, bh1]:  L += q0 bl0 + q1 bl1 (dp2a.lo, unsigned bytes), H += q0 bh0 + q1 bh1 (dp2a.hi, signed)
__device__ __forceinline__ void dot2_limbs(uint32_t q, uint32_t b, int32_t &H, uint32_t &L) {
  const uint32_t p = __byte_perm(b, 0u, 0x3120);
  asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(L) : "r"(q), "r"(p));
  asm("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(H) : "r"(q), "r"(p));
}

// Re-rank: one warp per source tile.  pair_feat = features of dictionary tile t recoloured with palette p,
// laid out [n_dict][n_pal][192] (built once per dictionary by features_i16_kernel<2>).
__global__ void __launch_bounds__(128)
match_rerank_kernel(const int16_t *__restrict__ q_feat, int64_t n_q, const int32_t *__restrict__ knn_idx, int k,
                    const int32_t *__restrict__ dict_pal, int64_t n_dict, int n_pal, const int16_t *__restrict__ pair_feat,
                    const uint32_t *__restrict__ pair_norm, int32_t *__restrict__ out_tile, int32_t *__restrict__ out_pal,
                    uint32_t *__restrict__ out_err) {
  __shared__ uint32_t s_q[4][96];
  __shared__ int32_t s_tile[4][64];
  __shared__ int32_t s_pal[4][64];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * 4 + w;
  if (qi >= n_q) return;
  const uint32_t *pq = reinterpret_cast<const uint32_t *>(q_feat + qi * 192);
  for (int j = lane; j < 96; j += 32) s_q[w][j] = __ldg(pq + j);
  // candidate tiles and their palettes; invalid slots (dictionary smaller than k) are dropped
  int n_t = 0, n_p = 0;
  for (int base = 0; base < k; base += 32) {
    const int j = base + lane;
    int32_t t = j < k ? __ldg(knn_idx + qi * k + j) : -1;
    if (t < 0 || t >= n_dict) t = -1;
    const int32_t p = t >= 0 ? __ldg(dict_pal + t) : -1;
    // tiles returned by the k-NN are distinct; palettes are not: keep the first occurrence of each
    const uint32_t tm = __ballot_sync(0xffffffffu, t >= 0);
    if (t >= 0) s_tile[w][n_t + __popc(tm & ((1u << lane) - 1u))] = t;
    n_t += __popc(tm);
    bool fresh = p >= 0;
    __syncwarp();
    for (int e = 0; e < n_p; ++e) fresh = fresh && (s_pal[w][e] != p);
    const uint32_t same = __match_any_sync(0xffffffffu, p);
    fresh = fresh && ((same & ((1u << lane) - 1u)) == 0);  // lowest lane holding this palette
    const uint32_t pm = __ballot_sync(0xffffffffu, fresh);
    if (fresh) s_pal[w][n_p + __popc(pm & ((1u << lane) - 1u))] = p;
    n_p += __popc(pm);
    __syncwarp();
  }
  // lexicographic minimum of (err, tile, pal) == strict-minimum scan in ascending (tile, pal) order.
  // Eight lanes share one candidate row: lane g of the group reads 48 contiguous bytes, so a warp-wide load touches four
  // 384-byte rows in whole 128-byte lines (one lane per row meant 32 different lines per load instruction and made the kernel
  // wait on L1/L2 wavefronts); the 8 partial sums are folded with three shuffles.  Every lane of a group ends up with the same
  // running best, so the final warp reduction is unchanged.
  // The distance is |q|^2 + |b|^2 - 2 q.b (mod 2^32, the reference's Cardinal accumulator): |b|^2 comes from a table built with the
  // pair features, q.b from the byte dot-product unit -- one permute and two dp2a per coefficient pair instead of two sign
  // extensions, two subtractions and two multiply-adds per coefficient.
  uint32_t best_e = 0xFFFFFFFFu;
  int32_t best_t = 0x7FFFFFFF, best_p = 0x7FFFFFFF;
  const int total = n_t * n_p;
  const int grp = lane >> 3, gl = lane & 7;
  uint32_t qreg[12];                                   // this lane's 24 query coefficients: words [12 gl, 12 gl + 12)
  uint32_t nq = 0;
#pragma unroll
  for (int c = 0; c < 12; ++c) { qreg[c] = s_q[w][12 * gl + c]; nq += sqdiff2(qreg[c], 0u); }
  nq += __shfl_xor_sync(0xffffffffu, nq, 1);
  nq += __shfl_xor_sync(0xffffffffu, nq, 2);
  nq += __shfl_xor_sync(0xffffffffu, nq, 4);
  for (int e0 = 0; e0 < total; e0 += 4) {
    const int e = e0 + grp;
    const bool live = e < total;
    const int32_t t = live ? s_tile[w][e / n_p] : 0, p = live ? s_pal[w][e % n_p] : 0;
    uint32_t dot = 0, nb = 0;
    if (live) {
      const size_t prow = (size_t)t * n_pal + p;
      const uint4 *pf = reinterpret_cast<const uint4 *>(pair_feat + prow * 192) + 3 * gl;
      nb = __ldg(pair_norm + prow);
      int32_t H = 0;
      uint32_t L = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 f = __ldg(pf + c);
        dot2_limbs(qreg[4 * c], f.x, H, L); dot2_limbs(qreg[4 * c + 1], f.y, H, L);
        dot2_limbs(qreg[4 * c + 2], f.z, H, L); dot2_limbs(qreg[4 * c + 3], f.w, H, L);
      }
      dot = ((uint32_t)H << 8) + L;
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    const uint32_t s = nq + nb - (dot << 1);
    if (live && (s < best_e || (s == best_e && (t < best_t || (t == best_t && p < best_p))))) { best_e = s; best_t = t; best_p = p; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint32_t oe = __shfl_xor_sync(0xffffffffu, best_e, o);
    const int32_t ot = __shfl_xor_sync(0xffffffffu, best_t, o), op = __shfl_xor_sync(0xffffffffu, best_p, o);
    if (oe < best_e || (oe == best_e && (ot < best_t || (ot == best_t && op < best_p)))) { best_e = oe; best_t = ot; best_p = op; }
  }
  if (lane == 0) {
    // the reference only accepts err < High(Cardinal) (knnErr starts at High(Cardinal), strict <, :1579-1597)
    const bool ok = total > 0 && best_e != 0xFFFFFFFFu;
    out_tile[qi] = ok ? best_t : -1;
    out_pal[qi] = ok ? best_p : -1;
    out_err[qi] = ok ? best_e : 0xFFFFFFFFu;
  }
}

// Exact NN of f64 vectors, summation order of CompareEuclidean (utils.pas:727-734): lane = query, 4 warps split the
// dictionary, first minimum in dictionary order.
constexpr int F64_Q = 32;
__global__ void __launch_bounds__(128)
knn_f64_kernel(const double *__restrict__ dict, int64_t n_dict, int dim, const double *__restrict__ q, int64_t n_q,
               int32_t *__restrict__ idx, double *__restrict__ dist) {
  extern __shared__ double s_qv[];  // [F64_Q][dim+1]
  __shared__ double s_best[4][F64_Q];
  __shared__ int32_t s_bi[4][F64_Q];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = (int64_t)blockIdx.x * F64_Q;
  const int ld = dim + 1;
  for (int i = threadIdx.x; i < F64_Q * dim; i += 128) {
    const int r = i / dim, c = i % dim;
    s_qv[r * ld + c] = (q0 + r < n_q) ? q[(q0 + r) * dim + c] : 0.0;
  }
  __syncthreads();
  const double *qv = s_qv + lane * ld;
  const int64_t per = (n_dict + 3) / 4, lo = w * per, hi = (lo + per < n_dict) ? lo + per : n_dict;
  double best = INFINITY;
  int32_t bi = -1;
  for (int64_t d = lo; d < hi; ++d) {
    const double *dv = dict + d * dim;
    double s = 0.0;
    for (int j = 0; j < dim; ++j) {
      const double df = __dsub_rn(qv[j], __ldg(dv + j));
      s = __dadd_rn(s, __dmul_rn(df, df));
    }
    if (s < best) { best = s; bi = (int32_t)d; }
  }
  s_best[w][lane] = best;
  s_bi[w][lane] = bi;
  __syncthreads();
  if (w == 0 && q0 + lane < n_q) {
    for (int o = 1; o < 4; ++o)
      if (s_best[o][lane] < best) { best = s_best[o][lane]; bi = s_bi[o][lane]; }  // strict: earlier range wins ties
    idx[q0 + lane] = bi;
    if (dist) dist[q0 + lane] = best;
  }
}

__global__ void match_plain_kernel(const int32_t *__restrict__ knn_idx, const uint32_t *__restrict__ knn_dist, int64_t n_q,
                                   const int32_t *__restrict__ dict_pal, int64_t n_dict, int32_t *__restrict__ out_tile,
                                   int32_t *__restrict__ out_pal, uint32_t *__restrict__ out_err) {
  // non-extended mode (tilingencoder.pas:1542-1558): nearest tile with its own palette
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q) return;
  const int32_t t = knn_idx[i];
  const bool ok = t >= 0 && t < n_dict;
  out_tile[i] = ok ? t : -1;
  out_pal[i] = ok ? dict_pal[t] : -1;
  out_err[i] = ok ? knn_dist[i] : 0xFFFFFFFFu;
}

int launch_match_plain(const int32_t *knn_idx, const uint32_t *knn_dist, int64_t n_q, const int32_t *dict_pal, int64_t n_dict,
                       int32_t *out_tile, int32_t *out_pal, uint32_t *out_err, cudaStream_t st) {
  if (n_q <= 0) return TM_OK;
  match_plain_kernel<<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(knn_idx, knn_dist, n_q, dict_pal, n_dict, out_tile, out_pal, out_err);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// Mirror-variant search (north star: "including H/V-mirror ... variants"; BASELINE configs[4]): the match of a source tile is
// tried in its stored orientation and H-, V-, HV-mirrored; variant v's result arrays are [v][n].  The FIRST strict minimum of
// the error over v = 0..3 wins, so the canonical orientation keeps ties (the reference never searches mirrors: v = 0 is its
// result, tilingencoder.pas:1393-1411, 1625-1637).
__global__ void __launch_bounds__(256) mirror_combine_kernel(const int32_t *__restrict__ t4, const int32_t *__restrict__ p4,
                                                             const uint32_t *__restrict__ e4, int64_t n, int32_t *__restrict__ out_tile,
                                                             int32_t *__restrict__ out_pal, uint32_t *__restrict__ out_err,
                                                             uint8_t *__restrict__ out_variant) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int bv = 0;
  uint32_t be = e4[i];
#pragma unroll
  for (int v = 1; v < 4; ++v) {
    const uint32_t e = e4[(int64_t)v * n + i];
    if (e < be) { be = e; bv = v; }
  }
  out_tile[i] = t4[(int64_t)bv * n + i];
  out_pal[i] = p4[(int64_t)bv * n + i];
  out_err[i] = be;
  out_variant[i] = (uint8_t)bv;
}
int launch_mirror_combine(const int32_t *t4, const int32_t *p4, const uint32_t *e4, int64_t n, int32_t *out_tile, int32_t *out_pal,
                          uint32_t *out_err, uint8_t *out_variant, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  mirror_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t4, p4, e4, n, out_tile, out_pal, out_err, out_variant);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_distance_pairs(const int16_t *a, const int16_t *b, int64_t n, uint32_t *out, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  distance_pairs_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(a, b, n, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_match_rerank(const int16_t *q_feat, int64_t n_q, const int32_t *knn_idx, int k, const int32_t *dict_pal,
                        const uint8_t *dict_idx, int64_t n_dict, const int32_t *palettes, int pal_size, int n_pal,
                        const int16_t *pair_feat, const uint32_t *pair_norm, int32_t *out_tile, int32_t *out_pal, uint32_t *out_err,
                        cudaStream_t st) {
  (void)dict_idx; (void)palettes; (void)pal_size;
  if (n_q <= 0) return TM_OK;
  if (k < 1 || k > 64 || !pair_feat || !pair_norm) return TM_ERR_ARG;
  ProfScope prof("rerank", st);
  match_rerank_kernel<<<(unsigned)((n_q + 3) / 4), 128, 0, st>>>(q_feat, n_q, knn_idx, k, dict_pal, n_dict, n_pal, pair_feat, pair_norm,
                                                                 out_tile, out_pal, out_err);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_row_norms(const int16_t *rows, int64_t n, uint32_t *norms, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  row_norms_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(rows, n, norms);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_knn_f64(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q, int32_t *idx, double *dist,
                   cudaStream_t st) {
  if (n_q <= 0) return TM_OK;
  if (dim < 1 || dim > 1024 || n_dict <= 0) return TM_ERR_ARG;
  const size_t smem = (size_t)F64_Q * (dim + 1) * sizeof(double);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(knn_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TM_ERR_CUDA;
  }
  knn_f64_kernel<<<(unsigned)((n_q + F64_Q - 1) / F64_Q), 128, smem, st>>>(dict, n_dict, dim, q, n_q, idx, dist);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
