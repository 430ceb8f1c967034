// dlquant.cu -- Dennis Lee's DL3 and DL1 colour quantisers as the reference carries them (dlquant/quantizer.c),
// bit-exact against that C code compiled with MSVC's type widths (`ulong` = 32 bits).
//
//   dl3quant (quantizer.c:437-455): build_table3 :480-508, calc_err :512-541, recount_next :543-560,
//     recount_dist :562-581, reduce_table3 :583-648, set_palette3 :650-663.
//     Colours are bucketed into 2^(3*bpc) cells; the pair with the least merge error is merged until quant_to remain.
//   dl1quant (quantizer.c:135-166): build_table1 :229-275, fixheap :277-300, reduce_table1 :302-322, set_palette1
//     :324-346.  5-bit-per-channel interleaved octree, a min-heap on cube population folds leaves into parents.
//
// GPU mapping.  The histogram is a parallel atomic pass over pixels.  DL3's pass 1 (nearest partner of every cell,
// O(C^2) error evaluations) is spread over the whole block; its pass 2 is C - quant_to strictly sequential merges,
// each followed by O(C) repairs: one 1024-thread block per image runs that loop with block-wide arg-min reductions
// and warp-cooperative partner rescans, all images of a batch (one per palette) in parallel.  DL1's heap loop is a
// chain of dependent O(log C) sifts: one thread per image, images in parallel.  Float error terms use explicitly
// rounded operations (no FMA contraction) so comparisons resolve exactly as in the C code.
#include "tm_kernels.h"
#include <math.h>

namespace tmg {

// ------------------------------------------------------------------ shared: per-image histograms
struct Dl3Cell { unsigned int r, g, b, n; };   // sums are `ulong` = 32-bit in the DLL build

__global__ void dl3_hist_kernel(const uint8_t *__restrict__ rgb, const int64_t *__restrict__ img_off, int n_img, int bpc,
                                Dl3Cell *__restrict__ cells /* [n_img][1 << 3*bpc] */) {
  const int img = blockIdx.y;
  const int64_t p0 = img_off[img], p1 = img_off[img + 1];
  const int mbpc = (1 << bpc) - 1;
  Dl3Cell *c = cells + ((int64_t)img << (3 * bpc));
  for (int64_t p = p0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += (int64_t)gridDim.x * blockDim.x) {
    const int R = rgb[p * 3], G = rgb[p * 3 + 1], B = rgb[p * 3 + 2];
    const int idx = (B * mbpc / 255) | ((G * mbpc / 255) << bpc) | ((R * mbpc / 255) << (2 * bpc));
    atomicAdd(&c[idx].r, (unsigned)R);
    atomicAdd(&c[idx].g, (unsigned)G);
    atomicAdd(&c[idx].b, (unsigned)B);
    atomicAdd(&c[idx].n, 1u);
  }
}

// ------------------------------------------------------------------ DL3
struct Dl3Tab {   // SoA view of one image's compacted table
  uint4 *sum;     // r, g, b, pixel_count
  unsigned int *rgb8;   // rr | gg << 8 | bb << 16
  float *err;
  int *cc;
};

__device__ __forceinline__ unsigned int dl3_setrgb(const uint4 s) {   // setrgb, quantizer.c:472-478
  const unsigned int v = s.w, v2 = (unsigned int)((int)v >> 1);
  const unsigned int rr = (s.x + v2) / v, gg = (s.y + v2) / v, bb = (s.z + v2) / v;
  return (rr & 255u) | ((gg & 255u) << 8) | ((bb & 255u) << 16);
}

__device__ __forceinline__ float dl3_calc_err(const uint4 a, const unsigned int a8, const uint4 b, const unsigned int b8) {
  // calc_err, quantizer.c:512-541
  const unsigned int P1 = a.w, P2 = b.w, P3 = P1 + P2;
  const int R3 = (int)((a.x + b.x + (P3 >> 1)) / P3), G3 = (int)((a.y + b.y + (P3 >> 1)) / P3), B3 = (int)((a.z + b.z + (P3 >> 1)) / P3);
  const int R1 = a8 & 255, G1 = (a8 >> 8) & 255, B1 = (a8 >> 16) & 255;
  const int R2 = b8 & 255, G2 = (b8 >> 8) & 255, B2 = (b8 >> 16) & 255;
  const float s1 = __fadd_rn(__fadd_rn((float)((R3 - R1) * (R3 - R1)), (float)((G3 - G1) * (G3 - G1))), (float)((B3 - B1) * (B3 - B1)));
  const float d1 = __fmul_rn(__fsqrt_rn(s1), (float)P1);
  const float s2 = __fadd_rn(__fadd_rn((float)((R2 - R3) * (R2 - R3)), (float)((G2 - G3) * (G2 - G3))), (float)((B2 - B3) * (B2 - B3)));
  const float d2 = __fmul_rn(__fsqrt_rn(s2), (float)P2);
  return __fadd_rn(d1, d2);
}

// recount_next (quantizer.c:543-560) by one warp: first minimum over j in (i, tot)
__device__ void dl3_recount_next_warp(const Dl3Tab &t, int i, int tot, int lane) {
  const uint4 a = t.sum[i];
  const unsigned int a8 = t.rgb8[i];
  float best = INFINITY;
  int bj = 0x7fffffff;
  for (int j = i + 1 + lane; j < tot; j += 32) {
    const float e = dl3_calc_err(a, a8, t.sum[j], t.rgb8[j]);
    if (e < best) { best = e; bj = j; }   // ascending j per lane: first minimum per lane
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float oe = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (oe < best || (oe == best && oj < bj)) { best = oe; bj = oj; }
  }
  if (lane == 0) {
    t.err[i] = best;
    t.cc[i] = (best < INFINITY) ? bj : 0;   // c2 starts at 0 and only moves on a strict improvement
  }
}

constexpr int DL3_T = 1024;

__global__ void __launch_bounds__(DL3_T)
dl3_reduce_kernel(const Dl3Cell *__restrict__ cells, int bpc, int quant_to, uint4 *__restrict__ sum_all, unsigned int *__restrict__ rgb8_all,
                  float *__restrict__ err_all, int *__restrict__ cc_all, int *__restrict__ list_all, uint8_t *__restrict__ pal_out,
                  int32_t *__restrict__ count_out) {
  const int img = blockIdx.x;
  const int C = 1 << (3 * bpc);
  const Dl3Cell *cell = cells + (int64_t)img * C;
  Dl3Tab t;
  t.sum = sum_all + (int64_t)img * C;
  t.rgb8 = rgb8_all + (int64_t)img * C;
  t.err = err_all + (int64_t)img * C;
  t.cc = cc_all + (int64_t)img * C;
  int *list = list_all + (int64_t)img * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_scan[DL3_T];
  __shared__ int s_tot, s_c1, s_c2, s_nlist;
  __shared__ float s_rerr[32];
  __shared__ int s_ridx[32];

  // ---- build_table3 tail: compact the non-empty cells in ascending cell order (quantizer.c:500-507)
  const int per = (C + DL3_T - 1) / DL3_T;
  int mine = 0;
  for (int i = tid * per; i < min(C, (tid + 1) * per); ++i) mine += cell[i].n != 0;
  s_scan[tid] = mine;
  __syncthreads();
  for (int o = 1; o < DL3_T; o <<= 1) {
    const int v = tid >= o ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  int pos = s_scan[tid] - mine;
  for (int i = tid * per; i < min(C, (tid + 1) * per); ++i) {
    const Dl3Cell c = cell[i];
    if (c.n) {
      const uint4 s = make_uint4(c.r, c.g, c.b, c.n);
      t.sum[pos] = s;
      t.rgb8[pos] = dl3_setrgb(s);
      ++pos;
    }
  }
  if (tid == DL3_T - 1) { s_tot = s_scan[tid]; s_c1 = 0; }
  __syncthreads();
  int tot = s_tot;

  // ---- pass 1 (quantizer.c:589-599): nearest partner of every entry but the last
  for (int i = warp; i < tot - 1; i += DL3_T / 32) dl3_recount_next_warp(t, i, tot, lane);
  if (tid == 0 && tot > 0) { t.err[tot - 1] = INFINITY; t.cc[tot - 1] = tot; }
  __syncthreads();

  // ---- pass 2 (quantizer.c:603-643)
  while (tot > quant_to) {
    // c1 = first index of the smallest err (strict < scan from HUGE_VALF; keeps the previous c1 if nothing is finite)
    float be = INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < tot; i += DL3_T) {
      const float e = t.err[i];
      if (e < be) { be = e; bi = i; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float oe = __shfl_xor_sync(0xffffffffu, be, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oe < be || (oe == be && oi < bi)) { be = oe; bi = oi; }
    }
    if (lane == 0) { s_rerr[warp] = be; s_ridx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      be = s_rerr[lane]; bi = s_ridx[lane];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float oe = __shfl_xor_sync(0xffffffffu, be, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oe < be || (oe == be && oi < bi)) { be = oe; bi = oi; }
      }
      if (lane == 0) {
        const int c1 = (be < INFINITY) ? bi : s_c1;
        const int c2 = t.cc[c1];
        // merge c1 into c2, shrink, move the last entry into slot c1, re-arm the new last entry (:617-627)
        uint4 a = t.sum[c1], b = t.sum[c2];
        b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
        t.sum[c2] = b;
        t.rgb8[c2] = dl3_setrgb(b);
        const int nt = tot - 1;
        t.sum[c1] = t.sum[nt]; t.rgb8[c1] = t.rgb8[nt]; t.err[c1] = t.err[nt]; t.cc[c1] = t.cc[nt];
        t.err[nt - 1] = INFINITY;
        t.cc[nt - 1] = nt;
        s_c1 = c1; s_c2 = c2; s_tot = nt; s_nlist = 0;
      }
    }
    __syncthreads();
    tot = s_tot;
    const int c1 = s_c1, c2 = s_c2;
    // (:629-637) entries that pointed at the moved entry: retarget below c1, rescan above c1
    for (int i = tid; i < tot; i += DL3_T) {
      if (t.cc[i] == tot) {
        if (i < c1) t.cc[i] = c1;
        else if (i > c1) list[atomicAdd(&s_nlist, 1)] = i;
      }
    }
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {   // recount_dist(c1), then recount_dist(c2) if c2 != tot (:639-640)
      const int c = rep == 0 ? c1 : c2;
      if (rep == 1 && c2 == tot) break;
      if (tid == 0) list[atomicAdd(&s_nlist, 1)] = c;
      const uint4 cs = t.sum[c];
      const unsigned int c8 = t.rgb8[c];
      for (int i = tid; i < c; i += DL3_T) {
        if (t.cc[i] == c) list[atomicAdd(&s_nlist, 1)] = i;
        else {
          const float e = dl3_calc_err(t.sum[i], t.rgb8[i], cs, c8);
          if (e < t.err[i]) { t.err[i] = e; t.cc[i] = c; }
        }
      }
      __syncthreads();
      const int nl = s_nlist;
      for (int li = warp; li < nl; li += DL3_T / 32) dl3_recount_next_warp(t, list[li], tot, lane);
      __syncthreads();
      if (tid == 0) s_nlist = 0;
      __syncthreads();
    }
  }

  // ---- set_palette3 (:650-663)
  for (int i = tid; i < quant_to; i += DL3_T) {
    const unsigned int v = i < tot ? t.rgb8[i] : 0u;
    pal_out[((int64_t)img * quant_to + i) * 3 + 0] = (uint8_t)(v & 255);
    pal_out[((int64_t)img * quant_to + i) * 3 + 1] = (uint8_t)((v >> 8) & 255);
    pal_out[((int64_t)img * quant_to + i) * 3 + 2] = (uint8_t)((v >> 16) & 255);
  }
  if (tid == 0 && count_out) count_out[img] = tot < quant_to ? tot : quant_to;
}

// ------------------------------------------------------------------ DL1
struct Dl1Cube { unsigned int r, g, b, pixel_count, pixels_in_cube; unsigned char children, pad[3]; };
constexpr int DL1_CUBES = 1 + 8 + 64 + 512 + 4096 + 32768;
__device__ __constant__ int c_dl1_base[6] = {0, 1, 9, 73, 585, 4681};

__global__ void dl1_hist_kernel(const uint8_t *__restrict__ rgb, const int64_t *__restrict__ img_off, Dl1Cube *__restrict__ cubes) {
  const int img = blockIdx.y;
  const int64_t p0 = img_off[img], p1 = img_off[img + 1];
  Dl1Cube *leaf = cubes + (int64_t)img * DL1_CUBES + 4681;
  for (int64_t p = p0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += (int64_t)gridDim.x * blockDim.x) {
    const int R = rgb[p * 3], G = rgb[p * 3 + 1], B = rgb[p * 3 + 2];
    // r_offset / g_offset / b_offset, quantizer.c:183-189: bit-interleaved 5-bit octree index
    const int ro = (R & 128) << 7 | (R & 64) << 5 | (R & 32) << 3 | (R & 16) << 1 | (R & 8) >> 1;
    const int go = (G & 128) << 6 | (G & 64) << 4 | (G & 32) << 2 | (G & 16) << 0 | (G & 8) >> 2;
    const int bo = (B & 128) << 5 | (B & 64) << 3 | (B & 32) << 1 | (B & 16) >> 1 | (B & 8) >> 3;
    Dl1Cube *c = leaf + (ro + go + bo);
    atomicAdd(&c->r, (unsigned)R);
    atomicAdd(&c->g, (unsigned)G);
    atomicAdd(&c->b, (unsigned)B);
    atomicAdd(&c->pixel_count, 1u);
  }
}

// one thread per image: build_table1's heap construction, reduce_table1, set_palette1 -- literally sequential
__global__ void dl1_reduce_kernel(Dl1Cube *__restrict__ cubes_all, unsigned int *__restrict__ heap_all, int lookup_size, int quant_to,
                                  int n_img, uint8_t *__restrict__ pal_out, int32_t *__restrict__ count_out) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  Dl1Cube *cube = cubes_all + (int64_t)img * DL1_CUBES;
  unsigned int *heap = heap_all + (int64_t)img * 32769;   // entry = level << 16 | index
#define CUBE(lv, ix) cube[c_dl1_base[lv] + (ix)]
#define HCUBE(h) CUBE((h) >> 16, (h) & 0xffff)
  int tot = 0;
  for (int i = 0; i < lookup_size; ++i) {
    const unsigned int cur = CUBE(5, i).pixel_count;
    if (cur) {
      heap[++tot] = (5u << 16) | (unsigned)i;
      CUBE(5, i).pixels_in_cube = cur;
      int head = i;
      for (int j = 4; j >= 0; --j) {
        const int tail = head & 7;
        head >>= 3;
        CUBE(j, head).pixels_in_cube += cur;
        CUBE(j, head).children |= (unsigned char)(1 << tail);
      }
    }
  }
  auto fixheap = [&](int id) {   // quantizer.c:277-300
    const unsigned int thres = heap[id];
    const unsigned int thres_val = HCUBE(thres).pixels_in_cube;
    const int half = tot >> 1;
    while (id <= half) {
      int index = id << 1;
      if (index < tot && HCUBE(heap[index]).pixels_in_cube > HCUBE(heap[index + 1]).pixels_in_cube) ++index;
      if (thres_val <= HCUBE(heap[index]).pixels_in_cube) break;
      heap[id] = heap[index];
      id = index;
    }
    heap[id] = thres;
  };
  for (int i = tot; i > 0; --i) fixheap(i);
  while (tot > quant_to) {   // reduce_table1, quantizer.c:302-322
    const unsigned int top = heap[1];
    const int tmp_level = top >> 16, tmp_index = top & 0xffff, t_level = tmp_level - 1, t_index = tmp_index >> 3;
    Dl1Cube &par = CUBE(t_level, t_index);
    const Dl1Cube &cur = CUBE(tmp_level, tmp_index);
    if (par.pixel_count) heap[1] = heap[tot--];
    else heap[1] = ((unsigned)t_level << 16) | (unsigned)t_index;
    par.pixel_count += cur.pixel_count;
    par.r += cur.r; par.g += cur.g; par.b += cur.b;
    par.children &= (unsigned char)~(1 << (tmp_index & 7));
    fixheap(1);
  }
  // set_palette1, quantizer.c:324-346: depth-first, children 7..0 before the node itself (explicit stack)
  int pal_index = 0;
  unsigned int stk[48];
  int sp = 0;
  stk[sp++] = 0u;   // (level << 16 | index), bit 31 = "children done"
  while (sp > 0) {
    const unsigned int e = stk[--sp];
    const int level = (e >> 16) & 0x7fff, index = e & 0xffff;
    const Dl1Cube &c = CUBE(level, index);
    if (!(e >> 31) && c.children) {
      stk[sp++] = e | 0x80000000u;
      for (int i = 0; i <= 7; ++i)   // pushed 0..7 so that 7 is visited first
        if (c.children & (1 << i)) stk[sp++] = ((unsigned)(level + 1) << 16) | (unsigned)((index << 3) + i);
      continue;
    }
    if (c.pixel_count) {
      const unsigned int sum = c.pixel_count;
      if (pal_index < quant_to) {
        pal_out[((int64_t)img * quant_to + pal_index) * 3 + 0] = (uint8_t)((c.r + (sum >> 1)) / sum);
        pal_out[((int64_t)img * quant_to + pal_index) * 3 + 1] = (uint8_t)((c.g + (sum >> 1)) / sum);
        pal_out[((int64_t)img * quant_to + pal_index) * 3 + 2] = (uint8_t)((c.b + (sum >> 1)) / sum);
      }
      ++pal_index;
    }
  }
  for (int i = pal_index; i < quant_to; ++i)
    for (int ch = 0; ch < 3; ++ch) pal_out[((int64_t)img * quant_to + i) * 3 + ch] = 0;
  if (count_out) count_out[img] = pal_index < quant_to ? pal_index : quant_to;
#undef CUBE
#undef HCUBE
}

// ------------------------------------------------------------------ launchers
// rgb888: concatenated images, img_off[n_img + 1] in pixels; palettes out [n_img][quant_to][3]
int run_dl3quant(const uint8_t *rgb, const int64_t *img_off, int n_img, int64_t max_pixels, int quant_to, int bpc, uint8_t *pal_out,
                 int32_t *count_out, cudaStream_t st) {
  if (n_img < 1 || bpc < 1 || bpc > 5 || quant_to < 1 || quant_to > 65536) return TM_ERR_ARG;
  const int C = 1 << (3 * bpc);
  Dl3Cell *cells = nullptr; uint4 *sum = nullptr; unsigned int *rgb8 = nullptr; float *err = nullptr; int *cc = nullptr, *list = nullptr;
  int rc = TM_OK;
  const size_t n = (size_t)n_img * C;
  if (cudaMallocAsync(&cells, n * sizeof(Dl3Cell), st) != cudaSuccess || cudaMallocAsync(&sum, n * sizeof(uint4), st) != cudaSuccess ||
      cudaMallocAsync(&rgb8, n * 4, st) != cudaSuccess || cudaMallocAsync(&err, n * 4, st) != cudaSuccess ||
      cudaMallocAsync(&cc, n * 4, st) != cudaSuccess || cudaMallocAsync(&list, n * 4, st) != cudaSuccess)
    rc = TM_ERR_NOMEM;
  if (rc == TM_OK) {
    cudaMemsetAsync(cells, 0, n * sizeof(Dl3Cell), st);
    int gx = (int)((max_pixels + 255) / 256);
    if (gx > 1024) gx = 1024;
    if (gx < 1) gx = 1;
    dl3_hist_kernel<<<dim3(gx, n_img), 256, 0, st>>>(rgb, img_off, n_img, bpc, cells);
    dl3_reduce_kernel<<<n_img, DL3_T, 0, st>>>(cells, bpc, quant_to, sum, rgb8, err, cc, list, pal_out, count_out);
    note_launch(2);
    if (cudaGetLastError() != cudaSuccess) rc = TM_ERR_CUDA;
  }
  cudaFreeAsync(cells, st); cudaFreeAsync(sum, st); cudaFreeAsync(rgb8, st); cudaFreeAsync(err, st); cudaFreeAsync(cc, st); cudaFreeAsync(list, st);
  return rc;
}

int run_dl1quant(const uint8_t *rgb, const int64_t *img_off, int n_img, int64_t max_pixels, int quant_to, int bpc, uint8_t *pal_out,
                 int32_t *count_out, cudaStream_t st) {
  if (n_img < 1 || bpc < 1 || bpc > 5 || quant_to < 1 || quant_to > 65536) return TM_ERR_ARG;
  Dl1Cube *cubes = nullptr;
  unsigned int *heap = nullptr;
  int rc = TM_OK;
  if (cudaMallocAsync(&cubes, (size_t)n_img * DL1_CUBES * sizeof(Dl1Cube), st) != cudaSuccess ||
      cudaMallocAsync(&heap, (size_t)n_img * 32769 * 4, st) != cudaSuccess)
    rc = TM_ERR_NOMEM;
  if (rc == TM_OK) {
    cudaMemsetAsync(cubes, 0, (size_t)n_img * DL1_CUBES * sizeof(Dl1Cube), st);
    int gx = (int)((max_pixels + 255) / 256);
    if (gx > 1024) gx = 1024;
    if (gx < 1) gx = 1;
    dl1_hist_kernel<<<dim3(gx, n_img), 256, 0, st>>>(rgb, img_off, cubes);
    dl1_reduce_kernel<<<(n_img + 31) / 32, 32, 0, st>>>(cubes, heap, 1 << (3 * bpc), quant_to, n_img, pal_out, count_out);
    note_launch(2);
    if (cudaGetLastError() != cudaSuccess) rc = TM_ERR_CUDA;
  }
  cudaFreeAsync(cubes, st); cudaFreeAsync(heap, st);
  return rc;
}

}  // namespace tmg
