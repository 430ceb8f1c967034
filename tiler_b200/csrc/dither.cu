// dither.cu -- arbitrary-palette positional dithering of 8x8 tiles (tilingencoder.pas:2268-2724).
//
//   PreparePlan                       :2268-2301   -> plan_kernel (one block per palette)
//   ColorCompare                      :2323-2337   -> int32 here: |t| <= 1724 so every term fits
//   DeviseBestMixingPlanThomasKnoll   :2565-2612   -> dither_kernel<true>
//   DeviseBestMixingPlanYliluoma      :2339-2563   -> dither_kernel<false> (the shipped SSE4.1 asm semantics)
//   DitherTile                        :2688-2724   -> mirror handling folded into the threshold-map lookup
//
// One thread per pixel (the 64-step error-feedback loop of a pixel is strictly sequential; pixels are independent),
// 128 threads = two (tile, palette) pairs per block, the pair's plan staged in shared memory as int4 {r,g,b,luma}.
// The final "sort the picks by luma and take entry map[y,x]" is an order statistic: with distinct lumas it is read
// off per-colour pick counts walked in luma order; when a palette has two colours of equal luma the reference's own
// non-stable QuickSort (extern.pas:370-418) decides the order, so that procedure is replayed literally.
#include "tm_kernels.h"

namespace tmg {

constexpr int MAXP = 256;
constexpr int32_t NULL_COLOR = (int32_t)0xffff00ff;  // cDitheringNullColor, utils.pas:45

__constant__ uint8_t c_dither_map[64] = {0,  48, 12, 60, 3,  51, 15, 63, 32, 16, 44, 28, 35, 19, 47, 31, 8,  56, 4,  52, 11, 59,
                                         7,  55, 40, 24, 36, 20, 43, 27, 39, 23, 2,  50, 14, 62, 1,  49, 13, 61, 34, 18, 46, 30,
                                         33, 17, 45, 29, 10, 58, 6,  54, 9,  57, 5,  53, 42, 26, 38, 22, 41, 25, 37, 21};

struct Plan {
  int32_t count;      // non-null colours
  int32_t has_dup;    // two colours share a luma
  int4 col[MAXP];     // r, g, b, LumaPal (299r+587g+114b)
  uint8_t remap[MAXP];
  uint8_t order[MAXP];  // compact indices by ascending luma (valid when !has_dup)
};

__global__ void __launch_bounds__(256) plan_kernel(const int32_t *__restrict__ palettes, int pal_size, Plan *__restrict__ plans) {
  Plan *pl = plans + blockIdx.x;
  const int32_t *pal = palettes + (int64_t)blockIdx.x * pal_size;
  __shared__ int s_cnt, s_dup;
  __shared__ int s_luma[MAXP];
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int i = 0; i < pal_size && i < MAXP; ++i) {
      const int32_t c = pal[i];
      if (c == NULL_COLOR) continue;
      const int r = c & 255, g = (c >> 8) & 255, b = (c >> 16) & 255;
      const int l = r * 299 + g * 587 + b * 114;
      pl->col[cnt] = make_int4(r, g, b, l);
      pl->remap[cnt] = (uint8_t)i;
      s_luma[cnt] = l;
      ++cnt;
    }
    s_cnt = cnt;
    s_dup = 0;
    pl->count = cnt;
  }
  __syncthreads();
  const int cnt = s_cnt;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    int rank = 0, dup = 0;
    const int l = s_luma[i];
    for (int j = 0; j < cnt; ++j) {
      rank += (s_luma[j] < l) || (s_luma[j] == l && j < i);
      dup |= (j != i && s_luma[j] == l);
    }
    pl->order[rank] = (uint8_t)i;
    if (dup) s_dup = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) pl->has_dup = s_dup;
}

// The reference's QuickSort (extern.pas:370-418) on a byte list keyed by luma; sub-ranges are disjoint, so an
// explicit stack visiting them in any order reproduces the recursive procedure exactly.
__device__ void quicksort_bytes(uint8_t *d, int n, const int4 *__restrict__ col) {
  if (n < 2) return;
  int16_t stk[128];
  int sp = 0;
  stk[sp++] = 0;
  stk[sp++] = (int16_t)(n - 1);
  while (sp > 0) {
    int last = stk[--sp], first = stk[--sp];
    if (last <= first) continue;
    for (;;) {
      int I = first, J = last, P = (first + last) >> 1;
      do {
        while (col[d[I]].w < col[d[P]].w) ++I;
        while (col[d[J]].w > col[d[P]].w) --J;
        if (I <= J) {
          const uint8_t t = d[J]; d[J] = d[I]; d[I] = t;
          if (P == I) P = J; else if (P == J) P = I;
          ++I; --J;
        }
      } while (I <= J);
      if (first < J && sp < 126) { stk[sp++] = (int16_t)first; stk[sp++] = (int16_t)J; }
      first = I;
      if (I >= last) break;
    }
  }
}

template <bool TK>
__global__ void __launch_bounds__(128)
dither_kernel(const int32_t *__restrict__ rgb, const uint8_t *__restrict__ mirror_flags, const int32_t *__restrict__ pair_tile,
              const int32_t *__restrict__ pair_pal, int64_t n_pairs, const Plan *__restrict__ plans, int y2_mixed,
              uint8_t *__restrict__ out_idx) {
  extern __shared__ uint8_t s_raw[];
  // per half-block (64 threads = one pair): int4 col[maxc]; then per-thread pick counters [maxc][128] bytes
  __shared__ int s_count[2], s_dup[2];
  const int half = threadIdx.x >> 6, px = threadIdx.x & 63;
  const int64_t pair = (int64_t)blockIdx.x * 2 + half;
  const bool active = pair < n_pairs;
  const Plan *pl = active ? plans + __ldg(pair_pal + pair) : plans;
  if (px == 0) { s_count[half] = active ? pl->count : 0; s_dup[half] = active ? pl->has_dup : 0; }
  __syncthreads();
  const int maxc = max(max(s_count[0], s_count[1]), 1);
  int4 *s_col = reinterpret_cast<int4 *>(s_raw) + half * maxc;
  uint8_t *s_cnt = s_raw + 2 * maxc * sizeof(int4);   // [maxc][128]
  const int cnt = s_count[half];
  for (int i = px; i < cnt; i += 64) s_col[i] = pl->col[i];
  for (int i = 0; i < maxc; ++i) s_cnt[i * 128 + threadIdx.x] = 0;
  __syncthreads();
  if (!active) return;

  const int64_t tile = pair_tile ? (int64_t)__ldg(pair_tile + pair) : pair;
  const int fl = mirror_flags ? mirror_flags[tile] : 0;
  const int sy = px >> 3, sx = px & 7;
  const int ny = (fl & 2) ? 7 - sy : sy, nx = (fl & 1) ? 7 - sx : sx;   // natural-orientation position
  int map_value = c_dither_map[(ny << 3) | nx];
  const int32_t c = __ldg(rgb + tile * 64 + px);
  const int r = c & 255, g = (c >> 8) & 255, b = (c >> 16) & 255;
  uint8_t result = 0;
  if (cnt == 0) { out_idx[pair * 64 + px] = 0; return; }

  uint8_t list[64];
  int n_list;
  if (TK) {
    int e0 = 0, e1 = 0, e2 = 0;
    for (int it = 0; it < 64; ++it) {
      const int t0 = r + (e0 * 9) / 100, t1 = g + (e1 * 9) / 100, t2 = b + (e2 * 9) / 100;
      const int luma1 = t0 * 299 + t1 * 587 + t2 * 114;
      int least = 0x7fffffff, chosen = 0;
      for (int i = 0; i < cnt; ++i) {
        const int4 p = s_col[i];
        const int dr = t0 - p.x, dg = t1 - p.y, db = t2 - p.z;
        const int ld = (luma1 - p.w) / 1000;
        const int pen = (dr * dr + dg * dg + db * db) * 13 + ((ld * ld) << 5);
        if (pen < least) { least = pen; chosen = i; }
      }
      if (s_dup[half]) list[it] = (uint8_t)chosen;
      s_cnt[chosen * 128 + threadIdx.x]++;
      const int4 p = s_col[chosen];
      e0 += r - p.x; e1 += g - p.y; e2 += b - p.z;
    }
    n_list = 64;
  } else {
    // Yliluoma: 32-bit lanes {r, g, b, luma/1000}; running mean via FVecInv[t] = 65536 div t
    const uint32_t tgt[4] = {(uint32_t)r, (uint32_t)g, (uint32_t)b, (uint32_t)((r * 299 + g * 587 + b * 114) / 1000)};
    uint32_t so_far[4] = {0, 0, 0, 0};
    int plan_count = 0;
    while (plan_count < y2_mixed) {
      const int max_test = plan_count == 0 ? 1 : plan_count;
      uint32_t least = 0xffffffffu;
      bool have = false;
      int chosen = 0, chosen_t = plan_count + 1;
      for (int i = 0; i < cnt; ++i) {
        const int4 p = s_col[i];
        uint32_t add[4] = {(uint32_t)p.x, (uint32_t)p.y, (uint32_t)p.z, (uint32_t)(p.w / 1000)};
        uint32_t sum[4] = {so_far[0], so_far[1], so_far[2], so_far[3]};
        for (int t = plan_count + 1; t <= plan_count + max_test; ++t) {
          const uint32_t inv = 65536u / (uint32_t)t;
          uint32_t pen = 0;
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            sum[l] += add[l];
            add[l] += 1;
            const uint32_t d = ((sum[l] * inv) >> 16) - tgt[l];
            pen += (d * d) * (l == 3 ? 32u : 13u);
          }
          if (!have || pen < least) { have = true; least = pen; chosen = i; chosen_t = t; }
        }
      }
      int amount = chosen_t - plan_count;
      if (amount > 256 - plan_count) amount = 256 - plan_count;
      const int4 p = s_col[chosen];
      for (int a = 0; a < amount && plan_count + a < 64; ++a) list[plan_count + a] = (uint8_t)chosen;
      plan_count += amount;
      so_far[0] += (uint32_t)p.x * amount; so_far[1] += (uint32_t)p.y * amount;
      so_far[2] += (uint32_t)p.z * amount; so_far[3] += (uint32_t)(p.w / 1000) * amount;
    }
    n_list = plan_count;            // <= 2*y2_mixed - 1 <= 31
    map_value = (map_value * n_list) >> 6;
  }

  if (TK && !s_dup[half]) {
    // rank map_value among the 64 picks ordered by luma == walk colours in luma order, accumulating pick counts
    int acc = 0;
    for (int o = 0; o < cnt; ++o) {
      const int ci = pl->order[o];
      acc += s_cnt[ci * 128 + threadIdx.x];
      if (acc > map_value) { result = pl->remap[ci]; break; }
    }
  } else {
    quicksort_bytes(list, n_list, s_col);
    result = pl->remap[list[map_value]];
  }
  out_idx[pair * 64 + px] = result;
}

static Plan *g_plans_dev[TM_MAX_DEVICES] = {};   // mixing plans of the palettes, per device
static int g_plans_cap_dev[TM_MAX_DEVICES] = {};

int launch_dither(const int32_t *rgb, const uint8_t *mirror_flags, const int32_t *pair_tile, const int32_t *pair_pal,
                  int64_t n_pairs, const int32_t *palettes, int pal_size, int n_pal, int use_tk, int y2_mixed, uint8_t *out_idx,
                  cudaStream_t st) {
  if (n_pairs <= 0) return TM_OK;
  if (pal_size < 1 || pal_size > MAXP || n_pal < 1 || y2_mixed < 1 || y2_mixed > 16) return TM_ERR_ARG;
  Plan *&g_plans = g_plans_dev[cur_device()];
  int &g_plans_cap = g_plans_cap_dev[cur_device()];
  if (n_pal > g_plans_cap) {
    if (g_plans) cudaFree(g_plans);
    g_plans = nullptr;
    g_plans_cap = 0;
    if (cudaMalloc(&g_plans, sizeof(Plan) * (size_t)n_pal) != cudaSuccess) return TM_ERR_NOMEM;
    g_plans_cap = n_pal;
  }
  plan_kernel<<<n_pal, 256, 0, st>>>(palettes, pal_size, g_plans);
  note_launch(2);
  const size_t smem = (size_t)2 * pal_size * sizeof(int4) + (size_t)pal_size * 128;
  const unsigned grid = (unsigned)((n_pairs + 1) / 2);
  static bool attr[TM_MAX_DEVICES] = {};
  if (first_use_on_device(attr)) {
    cudaFuncSetAttribute(dither_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAXP * 16 + MAXP * 128);
    cudaFuncSetAttribute(dither_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAXP * 16 + MAXP * 128);
  }
  if (use_tk)
    dither_kernel<true><<<grid, 128, smem, st>>>(rgb, mirror_flags, pair_tile, pair_pal, n_pairs, g_plans, y2_mixed, out_idx);
  else
    dither_kernel<false><<<grid, 128, smem, st>>>(rgb, mirror_flags, pair_tile, pair_pal, n_pairs, g_plans, y2_mixed, out_idx);
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
