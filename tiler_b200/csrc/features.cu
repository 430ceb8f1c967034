// features.cu -- tile -> 192-d psycho-visual feature vectors, and the load-time mirror canonicalisation.
//
//   int16 path  : ConvertToCpnPixels (YUV) + ComputeCpnPixelsPsyVisFeatures(pvsWeightedDCT)
//                 (tilingencoder.pas:3049-3131) with DCTInner_asm's exact summation order (utils.pas:874-1035):
//                 f32 products, f32 adds of lanes i/i+4 and i+8/i+12, then f64 accumulation in two lanes.
//                 Rounding of the weighted coefficient is half-to-even (Pascal Round) -> bit-exact vs the oracle.
//   f64 path    : ComputeTilePsyVisFeatures (tilingencoder.pas:3133-3182), LAB or YUV, any DCT mode.
//   mirrors     : GetTileHVMirrorHeuristics + H/VMirrorTile (tilingencoder.pas:4865-4878, 3257-3311, 1393-1411).
//
// Mapping: one thread per output coefficient (192 threads = one tile per pass), the 3x64 colour-plane samples of the
// tile broadcast from shared memory, the DCT basis held TRANSPOSED in shared memory ([pixel][coefficient]) so the 32
// threads of a warp read 32 consecutive words.  Blocks are persistent over tiles (grid = k * SM count).
#include "tm_kernels.h"
#include <math.h>

namespace tmg {

__constant__ uint8_t c_snake[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                                    41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                                    46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// cDCTWeights, utils.pas:72-97
__constant__ double c_weights[192] = {
    1.6193873005, 2.2901594831, 2.08509755623, 1.48366094411, 1.00227514334, 0.678296995242, 0.466224900598, 0.3265091542,
    2.2901594831, 1.94321815382, 2.04793073064, 1.68731108984, 1.2305666963, 0.868920337363, 0.61280991668, 0.436405793551,
    2.08509755623, 2.04793073064, 1.34329019223, 1.09205635862, 0.875748795257, 0.670882927016, 0.501731932449, 0.372504254596,
    1.48366094411, 1.68731108984, 1.09205635862, 0.772819797575, 0.605636379554, 0.48309405692, 0.380429446972, 0.295774038565,
    1.00227514334, 1.2305666963, 0.875748795257, 0.605636379554, 0.448996256676, 0.352889268808, 0.283006984131, 0.226951348204,
    0.678296995242, 0.868920337363, 0.670882927016, 0.48309405692, 0.352889268808, 0.27032073436, 0.215017739696, 0.17408067321,
    0.466224900598, 0.61280991668, 0.501731932449, 0.380429446972, 0.283006984131, 0.215017739696, 0.168869545842, 0.136153931001,
    0.3265091542, 0.436405793551, 0.372504254596, 0.295774038565, 0.226951348204, 0.17408067321, 0.136153931001, 0.109083846276,
    1.91113096927, 2.46074210438, 1.18284184739, 1.14982565193, 1.05017074788, 0.898018824055, 0.74725392039, 0.615105596242,
    2.46074210438, 1.58529308355, 1.21363250036, 1.38190029285, 1.33100189972, 1.17428548929, 0.996404342439, 0.830890433625,
    1.18284184739, 1.21363250036, 0.978712413627, 1.02624506078, 1.03145147362, 0.960060382087, 0.849823426169, 0.731221236837,
    1.14982565193, 1.38190029285, 1.02624506078, 0.861317501629, 0.801821139099, 0.751437590932, 0.685398513368, 0.608694761374,
    1.05017074788, 1.33100189972, 1.03145147362, 0.801821139099, 0.676555426187, 0.605503172737, 0.55002013668, 0.495804539034,
    0.898018824055, 1.17428548929, 0.960060382087, 0.751437590932, 0.605503172737, 0.514674450957, 0.454353482512, 0.407050308965,
    0.74725392039, 0.996404342439, 0.849823426169, 0.685398513368, 0.55002013668, 0.454353482512, 0.389234902883, 0.342353999733,
    0.615105596242, 0.830890433625, 0.731221236837, 0.608694761374, 0.495804539034, 0.407050308965, 0.342353999733, 0.295530605237,
    2.03871978502, 2.62502345193, 1.26180942886, 1.11019789803, 1.01397751469, 0.867069376285, 0.721500455585, 0.593906509971,
    2.62502345193, 1.69112867013, 1.17180569821, 1.3342742857, 1.28513006198, 1.13381474809, 0.962064122248, 0.802254508198,
    1.26180942886, 1.17180569821, 0.944981930573, 0.990876405848, 0.995903384143, 0.926972725286, 0.820534991409, 0.706020324706,
    1.11019789803, 1.3342742857, 0.990876405848, 0.831632933426, 0.77418706195, 0.725539939514, 0.661776842059, 0.587716619023,
    1.01397751469, 1.28513006198, 0.995903384143, 0.77418706195, 0.653238524286, 0.584635025748, 0.531064164893, 0.478717061273,
    0.867069376285, 1.13381474809, 0.926972725286, 0.725539939514, 0.584635025748, 0.496936637883, 0.438694579826, 0.393021669543,
    0.721500455585, 0.962064122248, 0.820534991409, 0.661776842059, 0.531064164893, 0.438694579826, 0.375820256136, 0.330555063063,
    0.593906509971, 0.802254508198, 0.706020324706, 0.587716619023, 0.478717061273, 0.393021669543, 0.330555063063, 0.285345396658};

// DCT bases built on the host exactly as InitLuts does (tilingencoder.pas:1703-1714): [special][v][u][y][x].
// Device copies are transposed to [special][pixel][coefficient].
static float *g_lutT_f32_dev[TM_MAX_DEVICES] = {};   // [2][64][64] per device
static double *g_lutT_f64_dev[TM_MAX_DEVICES] = {};  // [2][64][64] per device
#define g_lutT_f32 (g_lutT_f32_dev[cur_device()])
#define g_lutT_f64 (g_lutT_f64_dev[cur_device()])

int features_init(cudaStream_t st) {
  if (g_lutT_f32) return TM_OK;
  static float hf[2][64][64];
  static double hd[2][64][64];
  const double PI = 3.14159265358979323846;
  for (int v = 0; v < 8; ++v)
    for (int u = 0; u < 8; ++u)
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
          // cDCTUVRatio is a single-precision table (utils.pas:100)
          const float rf = (v == 0 && u == 0) ? 0.5f : ((v == 0 || u == 0) ? (float)sqrt(0.5) : 1.0f);
          const double r = (double)rf;
          const double d0 = cos((x + 0.5) * u * PI / 8.0) * cos((y + 0.5) * v * PI / 8.0) * r;
          const double d1 = cos((x + 0.5) * u * PI / 16.0) * cos((y + 0.5) * v * PI / 16.0) * r;
          hd[0][y * 8 + x][v * 8 + u] = d0;
          hd[1][y * 8 + x][v * 8 + u] = d1;
          hf[0][y * 8 + x][v * 8 + u] = (float)d0;
          hf[1][y * 8 + x][v * 8 + u] = (float)d1;
        }
  float *pf = nullptr;
  double *pd = nullptr;
  if (cudaMalloc(&pf, sizeof(hf)) != cudaSuccess) return TM_ERR_NOMEM;
  if (cudaMalloc(&pd, sizeof(hd)) != cudaSuccess) return TM_ERR_NOMEM;
  if (cudaMemcpyAsync(pf, hf, sizeof(hf), cudaMemcpyHostToDevice, st) != cudaSuccess) return TM_ERR_CUDA;
  if (cudaMemcpyAsync(pd, hd, sizeof(hd), cudaMemcpyHostToDevice, st) != cudaSuccess) return TM_ERR_CUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess) return TM_ERR_CUDA;
  g_lutT_f32 = pf;
  g_lutT_f64 = pd;
  return TM_OK;
}

// RGBToYUV, utils.pas:478-490: double evaluation, single storage
__device__ __forceinline__ void rgb_to_yuv(int r, int g, int b, float &y, float &u, float &v) {
  const double yd = __dadd_rn(__dadd_rn(__dmul_rn((double)r, 299.0 / 1000.0), __dmul_rn((double)g, 587.0 / 1000.0)),
                              __dmul_rn((double)b, 114.0 / 1000.0));
  y = (float)yd;
  u = (float)__dmul_rn(__dsub_rn((double)b, (double)y), 0.492);
  v = (float)__dmul_rn(__dsub_rn((double)r, (double)y), 0.877);
}

// RGBToLAB, utils.pas:374-410 (D50).  pow() is the CUDA libm one: agrees with the oracle's to ~1 ulp -> LAB features
// are compared with a tolerance (they only feed the palette clustering).
__device__ void rgb_to_lab(int ir, int ig, int ib, float &ol, float &oa, float &ob) {
  float r = (float)(ir / 255.0), g = (float)(ig / 255.0), b = (float)(ib / 255.0);
  r = (r > 0.04045) ? (float)pow(((double)r + 0.055) / 1.055, 2.4) : (float)((double)r / 12.92);
  g = (g > 0.04045) ? (float)pow(((double)g + 0.055) / 1.055, 2.4) : (float)((double)g / 12.92);
  b = (b > 0.04045) ? (float)pow(((double)b + 0.055) / 1.055, 2.4) : (float)((double)b / 12.92);
  float x = (float)(((double)r * 0.49000 + (double)g * 0.31000 + (double)b * 0.20000) / 0.17697);
  float y = (float)(((double)r * 0.17697 + (double)g * 0.81240 + (double)b * 0.01063) / 0.17697);
  float z = (float)(((double)r * 0.00000 + (double)g * 0.01000 + (double)b * 0.99000) / 0.17697);
  x = (float)((double)x * (1 / (96.6797 / 100)));
  y = (float)((double)y * (1 / (100.000 / 100)));
  z = (float)((double)z * (1 / (82.5188 / 100)));
  x = (x > 0.008856) ? (float)pow((double)x, 1.0 / 3) : (float)((7.787 * (double)x) + 16.0 / 116);
  y = (y > 0.008856) ? (float)pow((double)y, 1.0 / 3) : (float)((7.787 * (double)y) + 16.0 / 116);
  z = (z > 0.008856) ? (float)pow((double)z, 1.0 / 3) : (float)((7.787 * (double)z) + 16.0 / 116);
  ol = (float)((116 * (double)y) - 16);
  oa = (float)(500 * ((double)x - (double)y));
  ob = (float)(200 * ((double)y - (double)z));
}

// DCTInner_asm order for one coefficient of the plane s_c[64] (shared memory, read as broadcast float4); lut[64] is the
// coefficient's own column of the DCT basis, held in REGISTERS: a thread computes the same coefficient for every tile, and
// fetching the basis from shared memory for every tile (64 loads per thread and tile, 768 wavefronts per tile) made the
// kernel shared-memory-bandwidth bound at 3x its arithmetic time.
__device__ __forceinline__ double dct_inner(const float *__restrict__ s_c, const float (&lut)[64]) {
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    float p[16];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 c4 = *reinterpret_cast<const float4 *>(s_c + s * 16 + v * 4);
      p[v * 4 + 0] = __fmul_rn(c4.x, lut[s * 16 + v * 4 + 0]);
      p[v * 4 + 1] = __fmul_rn(c4.y, lut[s * 16 + v * 4 + 1]);
      p[v * 4 + 2] = __fmul_rn(c4.z, lut[s * 16 + v * 4 + 2]);
      p[v * 4 + 3] = __fmul_rn(c4.w, lut[s * 16 + v * 4 + 3]);
    }
    const float a0 = __fadd_rn(p[0], p[4]), a1 = __fadd_rn(p[1], p[5]), a2 = __fadd_rn(p[2], p[6]), a3 = __fadd_rn(p[3], p[7]);
    const float b0 = __fadd_rn(p[8], p[12]), b1 = __fadd_rn(p[9], p[13]), b2 = __fadd_rn(p[10], p[14]), b3 = __fadd_rn(p[11], p[15]);
    const double l0 = __dadd_rn(__dadd_rn((double)a0, (double)b0), __dadd_rn((double)a2, (double)b2));
    const double l1 = __dadd_rn(__dadd_rn((double)a1, (double)b1), __dadd_rn((double)a3, (double)b3));
    acc0 = __dadd_rn(acc0, l0);
    acc1 = __dadd_rn(acc1, l1);
  }
  return __dadd_rn(acc0, acc1);
}

#ifndef FEAT_NP
#define FEAT_NP 2
#endif
// MODE 0: RGB tiles; 1: palette indices with per-tile palette; 2: every (tile, palette) pair, item = tile*n_pal + pal;
// 3: RGB tiles read through their mirror flags (pal_idx = flags[n]: ConvertToCpnPixels with AHMirror / AVMirror, :3049-3101);
// 4: sliding window over a frame buffer (rgb = frame [h][fw], item = oy * pw + ox: DoDCTs, :1437-1462)
template <int MODE>
__global__ void __launch_bounds__(192, 3)
features_i16_kernel(const int32_t *__restrict__ rgb, const uint8_t *__restrict__ pal_idx, int n_pal_all,
                    const int32_t *__restrict__ sel_pal, const int32_t *__restrict__ palettes, int pal_size, int64_t n,
                    const float *__restrict__ lutT, int16_t *__restrict__ out, int fw = 0, int pw = 0) {
  constexpr int NP = FEAT_NP;   // tiles per block iteration
  __shared__ __align__(16) float s_cpn[2][NP][3][64];   // [buffer][tile of the group][plane][pixel]: one barrier per GROUP of tiles
  __shared__ __align__(16) int16_t s_out[2][NP][192];
  const int t = threadIdx.x;
  const int c = t >> 6, vu = t & 63;
  float lut[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) lut[i] = __ldg(lutT + i * 64 + vu);
  const double wgt = c_weights[t];
  const int dst = c * 64 + c_snake[vu];
  // Colour conversion is spread over all 192 threads: thread t converts component t / 64 of pixel t % 64 (the luma every
  // component needs is recomputed, 5 double operations) instead of 64 threads converting while 128 wait at the barrier.
  const int px = t & 63;
  auto fetch = [&](int64_t tile) -> int32_t {
    if (tile >= n) return 0;
    if (MODE == 0) return __ldg(rgb + tile * 64 + px);
    if (MODE == 3) {
      const int fl = __ldg(pal_idx + tile);
      const int x = (fl & 1) ? 7 - (px & 7) : (px & 7), y = (fl & 2) ? 7 - (px >> 3) : (px >> 3);
      return __ldg(rgb + tile * 64 + y * 8 + x);
    }
    if (MODE == 4) {
      const int oy = (int)tile / pw, ox = (int)tile - oy * pw;
      return __ldg(rgb + (int64_t)(oy + (px >> 3)) * fw + ox + (px & 7));
    }
    const int64_t src = (MODE == 2) ? tile / n_pal_all : tile;
    const int32_t p = (MODE == 2) ? (int32_t)(tile % n_pal_all) : __ldg(sel_pal + tile);
    return __ldg(palettes + (int64_t)p * pal_size + __ldg(pal_idx + src * 64 + px));
  };
  auto convert = [&](int32_t col) -> float {   // RGBToYUV (utils.pas:478-490), component c only (c is warp-uniform)
    const int r = col & 255, g = (col >> 8) & 255, b = (col >> 16) & 255;
    const float y = (float)__dadd_rn(__dadd_rn(__dmul_rn((double)r, 299.0 / 1000.0), __dmul_rn((double)g, 587.0 / 1000.0)),
                                     __dmul_rn((double)b, 114.0 / 1000.0));
    float val = y;
    if (c != 0) val = (float)__dmul_rn(__dsub_rn((double)(c == 1 ? b : r), (double)y), c == 1 ? 0.492 : 0.877);
    return val;
  };
  // A block works on GROUPS of NP consecutive tiles: NP independent 64-term sums per thread share the basis registers and
  // multiply the instruction-level parallelism, and the group's NP x 384 output bytes leave as coalesced 4-byte stores.
  // Software pipeline: the pixels of the NEXT group are requested before the current group's sums, so the global load
  // latency never sits between two barriers; the previous group's coefficients are stored while the current one is computed.
  const int64_t n_groups = (n + NP - 1) / NP;
  int64_t grp = blockIdx.x;
  int32_t col[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) col[k] = grp < n_groups ? fetch(NP * grp + k) : 0;
  int buf = 0;
  int64_t prev_grp = -1;
  auto flush = [&](int64_t pg, int b) {
    const int64_t t0 = NP * pg;
    uint32_t *o = reinterpret_cast<uint32_t *>(out + t0 * 192);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&s_out[b][0][0]);
#pragma unroll
    for (int w = t; w < NP * 96; w += 192)
      if (t0 + w / 96 < n) o[w] = src[w];   // words [96 k, 96 k + 96) = tile NP pg + k
  };
  for (; grp < n_groups; grp += gridDim.x, buf ^= 1) {
#pragma unroll
    for (int k = 0; k < NP; ++k) s_cpn[buf][k][c][px] = convert(col[k]);
    __syncthreads();   // planes of this group visible; coefficients of the previous group complete in s_out[buf ^ 1]
    const int64_t next = grp + gridDim.x;
    if (next < n_groups) {
#pragma unroll
      for (int k = 0; k < NP; ++k) col[k] = fetch(NP * next + k);
    }
    if (prev_grp >= 0) flush(prev_grp, buf ^ 1);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      double z = dct_inner(s_cpn[buf][k][c], lut);
      z = __dmul_rn(z, wgt);
      s_out[buf][k][dst] = (int16_t)__double2int_rn(z);
    }
    prev_grp = grp;
  }
  __syncthreads();
  if (prev_grp >= 0) flush(prev_grp, buf ^ 1);
}

// ComputeTilePsyVisFeatures: f64, sequential 64-term sums (DCTInner<PDouble>, utils.pas:782-872)
__global__ void __launch_bounds__(192, 2)
features_f64_kernel(const int32_t *__restrict__ rgb, int64_t n, int weighted, int use_lab, const double *__restrict__ lutT,
                    double *__restrict__ out) {
  extern __shared__ double s_lutd[];  // 4096 doubles
  __shared__ double s_cpn[3][64];
  for (int i = threadIdx.x; i < 4096; i += 192) s_lutd[i] = lutT[i];
  const int t = threadIdx.x;
  const int c = t >> 6, vu = t & 63;
  const double wgt = weighted ? c_weights[t] : 1.0;
  const int dst = c * 64 + c_snake[vu];
  for (int64_t tile = blockIdx.x; tile < n; tile += gridDim.x) {
    __syncthreads();
    if (t < 64) {
      const int32_t col = __ldg(rgb + tile * 64 + t);
      float y, u, v;
      if (use_lab) rgb_to_lab(col & 255, (col >> 8) & 255, (col >> 16) & 255, y, u, v);
      else rgb_to_yuv(col & 255, (col >> 8) & 255, (col >> 16) & 255, y, u, v);
      s_cpn[0][t] = (double)y; s_cpn[1][t] = (double)u; s_cpn[2][t] = (double)v;
    }
    __syncthreads();
    double z = 0.0;
#pragma unroll 8
    for (int i = 0; i < 64; ++i) z = __dadd_rn(z, __dmul_rn(s_cpn[c][i], s_lutd[i * 64 + vu]));
    if (weighted) z = __dmul_rn(z, wgt);
    out[tile * 192 + dst] = z;
  }
}

// ComputeTilePsyVisFeatures, Mode = pvsWavelets (tilingencoder.pas:3151-3158 over WaveletGS :2727-2762): three levels of the
// normalised Haar transform per colour plane (8x8, then the 4x4 and the 2x2 low-pass corner), f64, stored through the zig-zag
// table like the DCT modes.  Not on any default path (DitheringMode defaults to the weighted special DCT); block = one tile,
// thread = (plane, position); same operation order as the reference, so the result is bit-exact.
__global__ void __launch_bounds__(192, 2)
features_wavelet_f64_kernel(const int32_t *__restrict__ rgb, int64_t n, int use_lab, double *__restrict__ out) {
  __shared__ double s_cur[3][64], s_tmp[3][64];
  const int t = threadIdx.x;
  const int c = t >> 6, i = t & 63, y = i >> 3, x = i & 7;
  const double factor = 1.0 / sqrt(2.0);
  for (int64_t tile = blockIdx.x; tile < n; tile += gridDim.x) {
    __syncthreads();
    if (t < 64) {
      const int32_t col = __ldg(rgb + tile * 64 + t);
      float py, pu, pv;
      if (use_lab) rgb_to_lab(col & 255, (col >> 8) & 255, (col >> 16) & 255, py, pu, pv);
      else rgb_to_yuv(col & 255, (col >> 8) & 255, (col >> 16) & 255, py, pu, pv);
      s_cur[0][t] = (double)py; s_cur[1][t] = (double)pu; s_cur[2][t] = (double)pv;
    }
    __syncthreads();
    for (int d = 8; d >= 2; d >>= 1) {
      if (y < d && x < d / 2) {                       // rows: low pass left, high pass right
        const double a = s_cur[c][2 * x + 8 * y], b = s_cur[c][2 * x + 1 + 8 * y];
        s_tmp[c][x + 8 * y] = __dmul_rn(__dadd_rn(a, b), factor);
        s_tmp[c][x + d / 2 + 8 * y] = __dmul_rn(__dsub_rn(a, b), factor);
      }
      __syncthreads();
      if (x < d && y < d / 2) {                       // columns: low pass top, high pass bottom
        const double a = s_tmp[c][x + 8 * (2 * y)], b = s_tmp[c][x + 8 * (2 * y + 1)];
        s_cur[c][x + 8 * y] = __dmul_rn(__dadd_rn(a, b), factor);
        s_cur[c][x + 8 * (y + d / 2)] = __dmul_rn(__dsub_rn(a, b), factor);
      }
      __syncthreads();
    }
    out[tile * 192 + c * 64 + c_snake[i]] = s_cur[c][i];
  }
}

// One thread per tile: quadrant luma sums, flags, in-place flip (tile stays 256 bytes in registers/L1)
__global__ void __launch_bounds__(128) mirror_kernel(int32_t *__restrict__ rgb, int64_t n, uint8_t *__restrict__ flags) {
  const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= n) return;
  int32_t *p = rgb + tile * 64;
  int q[4] = {0, 0, 0, 0};
  for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 8; ++i) {
      const int32_t col = p[j * 8 + i];
      q[(j >> 2) * 2 + (i >> 2)] += (col & 255) * 299 + ((col >> 8) & 255) * 587 + ((col >> 16) & 255) * 114;
    }
  const bool hm = (q[0] + q[2]) < (q[1] + q[3]);
  const bool vm = (q[0] + q[1]) < (q[2] + q[3]);
  flags[tile] = (uint8_t)((hm ? 1 : 0) | (vm ? 2 : 0));
  if (hm)
    for (int j = 0; j < 8; ++j)
      for (int i = 0; i < 4; ++i) { const int32_t a = p[j * 8 + i]; p[j * 8 + i] = p[j * 8 + 7 - i]; p[j * 8 + 7 - i] = a; }
  if (vm)
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < 8; ++i) { const int32_t a = p[j * 8 + i]; p[j * 8 + i] = p[(7 - j) * 8 + i]; p[(7 - j) * 8 + i] = a; }
}

// ------------------------------------------------------------------ sliding-window features, FAST mode (separable, f64)
// DoDCTs (tilingencoder.pas:1437-1462) asks for the weighted DCT of the 8x8 window at EVERY pixel offset of a frame buffer.
// The bit-exact kernel above spends 64 f32 products, their pair adds and 32 f32->f64 conversions per coefficient to reproduce
// DCTInner_asm's summation order, and nothing is shared between overlapping windows.  The 2-D basis is separable
// (lut[v][u][y][x] = cos((x+1/2) u pi/8) cos((y+1/2) v pi/8) ratio[v][u], tilingencoder.pas:1709), so in exact arithmetic
//   row pass   R[y][ox][u]      = sum_x P[y][ox+x] cos((x+1/2) u pi/8)        -- shared by the 8 windows that contain the row segment
//   column pass F[oy][ox][v][u] = sum_j R[oy+j][ox][u] cos((j+1/2) v pi/8)
// is 8 + 36 f64 operations per coefficient column (the column pass split into its even and odd halves) instead of 512.  Everything runs in f64 (B200 keeps a 1:2 FP64 rate), which is
// MORE accurate than the reference's f32 products; the result differs from the reference only where the reference's own f32
// rounding noise (~3e-4 absolute) straddles a .5 rounding boundary: <= 1 LSB, ~5e-6 of the coefficients on the synthetic clip
// (contract: <= 1 LSB, <= 1e-3 of the coefficients; SURVEY "Parity contract").  Selected with tm_set_feature_mode(1); the
// bit-exact kernel stays the default.
//
// Thread = (plane, ox, u): it marches down the rows of its window column with the last 8 row-pass values in a register ring
// (the loop is unrolled over the 8 ring phases, so the ring is statically indexed and the column-pass cosines are constant-bank
// operands), and emits the 8 coefficients (v = 0..7) of window (y - 7, ox) at every row y.  A block = 8 window columns x 3
// planes x 8 u; its 8 x 192 coefficients per row leave through shared memory as one contiguous 3 KB store.
__constant__ double c_cos8[8][8];      // [k][i] = cos((i + 1/2) k pi / 8)
__constant__ double c_scale[192];      // [cpn][v][u] = (double)cDCTUVRatio[v][u] * cDCTWeights[cpn][v][u]
constexpr int SF_OXB = 8;              // window columns per block
constexpr int SF_SEG = 64;             // window rows per block
constexpr int SF_CH = 16;              // frame rows converted per chunk
constexpr int SF_ROWP = 200;           // int16 per staged window row (192 + 8 of padding)

// cos(m pi / 16), m = 1..7: every entry of the 8-point DCT basis cos((j + 1/2) v pi / 8) is +-1 or +- one of these
// (read from constant memory once per thread and kept in uniform registers: written as literals the compiler re-materialises
// them with two UMOVs in front of every use)
__constant__ double c_sfc[8] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476,
                                0.5555702330196023, 0.38268343236508984, 0.19509032201612833};

// Column pass of one window column: F[v] = sum_j w_j cos((j + 1/2) v pi / 8), w_j = ring[(j + PH + 1) & 7] (window row j).
// Even / odd split (the basis is symmetric in j for even v, antisymmetric for odd v): 14 additions + 22 multiply-adds instead of
// 64 multiply-adds, and only SEVEN distinct constants.  The direct form asked for 64 distinct f64 constants per phase; the
// compiler kept them in uniform registers, ran out, and issued more R2UR / spill moves than DFMAs (632 + 186 against 576 in the
// 8 unrolled phases).
template <int PH>
__device__ __forceinline__ void sf_cols(const double (&ring)[8], const double (&C)[8], double (&F)[8]) {
  const double SF_C1 = C[1], SF_C2 = C[2], SF_C3 = C[3], SF_C4 = C[4], SF_C5 = C[5], SF_C6 = C[6], SF_C7 = C[7];
#define SF_W(j) ring[((j) + PH + 1) & 7]
  const double s0 = SF_W(0) + SF_W(7), s1 = SF_W(1) + SF_W(6), s2 = SF_W(2) + SF_W(5), s3 = SF_W(3) + SF_W(4);
  const double d0 = SF_W(0) - SF_W(7), d1 = SF_W(1) - SF_W(6), d2 = SF_W(2) - SF_W(5), d3 = SF_W(3) - SF_W(4);
#undef SF_W
  const double a0 = s0 + s3, a1 = s1 + s2, b0 = s0 - s3, b1 = s1 - s2;
  F[0] = a0 + a1;
  F[4] = (a0 - a1) * SF_C4;
  F[2] = fma(b0, SF_C2, b1 * SF_C6);
  F[6] = fma(b0, SF_C6, -(b1 * SF_C2));
  F[1] = fma(d3, SF_C7, fma(d2, SF_C5, fma(d1, SF_C3, d0 * SF_C1)));
  F[3] = fma(d3, -SF_C5, fma(d2, -SF_C1, fma(d1, -SF_C7, d0 * SF_C3)));
  F[5] = fma(d3, SF_C3, fma(d2, SF_C7, fma(d1, -SF_C1, d0 * SF_C5)));
  F[7] = fma(d3, -SF_C1, fma(d2, SF_C3, fma(d1, -SF_C5, d0 * SF_C7)));
}

template <int PH, bool LIMBS>
__device__ __forceinline__ void sf_step(const double *__restrict__ prow /* plane row, this thread's first pixel */, const double (&cu)[8],
                                        const double (&sc)[8], const int (&dst)[8], double (&ring)[8], const double (&C)[8], bool emit,
                                        int16_t *__restrict__ o, uint32_t *__restrict__ s_norm /* this window's norm accumulator */) {
  double r = 0.0;
#pragma unroll
  for (int x = 0; x < 8; ++x) r = fma(prow[x], cu[x], r);
  ring[PH] = r;
  if (emit) {   // block-uniform
    double F[8];
    sf_cols<PH>(ring, C, F);
    if (!LIMBS) {
#pragma unroll
      for (int v = 0; v < 8; ++v) o[dst[v]] = (int16_t)__double2int_rn(__dmul_rn(F[v], sc[v]));
    } else {
      // the staged row is the LIMB row [hi(192) | lo(192)] of the window, and the squares of this thread's 8 coefficients are summed
      // over the 8 threads (u) of its (plane, window) by three shuffles: one shared-memory add per (plane, window)
      uint8_t *ob = reinterpret_cast<uint8_t *>(o);
      uint32_t acc = 0;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        // |coefficient| <= 13 212 for 8-bit pixels (sum of the basis magnitudes x weights), so the int16 narrowing of the
        // other path is the identity here and the bytes below are the two's-complement limbs of the same value
        const int e = __double2int_rn(__dmul_rn(F[v], sc[v]));
        ob[dst[v]] = (uint8_t)((uint32_t)e >> 8);
        ob[192 + dst[v]] = (uint8_t)e;
        acc += (uint32_t)(e * e);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if ((threadIdx.x & 7) == 0) atomicAdd(s_norm, acc);
    }
  }
}

// LIMBS = true writes what the tensor-core motion search consumes instead of the int16 features: the limb row
// [hi(192) | lo(192)] of every window at its PADDED position (row pitch pwp, a multiple of 64) and its squared norm mod 2^32;
// positions ox in [fw - 7, pwp) get zero rows and zero norms (grid.x covers pwp / 8 column blocks).  That is
// cand_limb_split_kernel folded into the producer: the 349 MB int16 intermediate of a 720p frame is never written or read.
template <bool LIMBS>
__device__ __forceinline__ void sf_body(const int32_t *__restrict__ frame, int fw, int fh, int16_t *__restrict__ out,
                                        uint8_t *__restrict__ limbs, uint32_t *__restrict__ norms, int pwp) {
  __shared__ double s_pl[3][SF_CH][SF_OXB + 8];          // converted planes of the current row chunk (15 of 16 columns used)
  __shared__ __align__(16) int16_t s_out[2][SF_OXB][SF_ROWP];   // rows padded to 400 bytes: the limb flush reads 8 rows at once
  __shared__ uint32_t s_nrm[2][SF_OXB];
  const int t = threadIdx.x;
  const int u = t & 7, oxl = (t >> 3) & 7, cpn = t >> 6;
  const int pw = fw - 7, ph_rows = fh - 7;
  const int ox0 = blockIdx.x * SF_OXB;
  const int oy0 = blockIdx.y * SF_SEG;
  const int oy1 = min(oy0 + SF_SEG, ph_rows);              // window rows [oy0, oy1)
  const int y_end = oy1 + 7;                               // frame rows [oy0, y_end)
  if (LIMBS) {
    if (t < 2 * SF_OXB) (&s_nrm[0][0])[t] = 0;
    if (ox0 >= pw) {   // a column block of padding only: zero rows, zero norms
      const int wdw = t / 24, seg = t % 24;
      for (int oy = oy0; oy < oy1; ++oy) {
        uint8_t *dst = limbs + ((int64_t)oy * pwp + ox0 + wdw) * 384 + seg * 8;
        *reinterpret_cast<uint2 *>(dst) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2 *>(dst + 192) = make_uint2(0u, 0u);
        if (t < SF_OXB) norms[(int64_t)oy * pwp + ox0 + t] = 0u;
      }
      return;
    }
  }
  double cu[8], sc[8], ring[8], C[8];
  int dst[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    C[i] = c_sfc[i];
    cu[i] = c_cos8[u][i];
    sc[i] = c_scale[cpn * 64 + i * 8 + u];
    dst[i] = cpn * 64 + c_snake[i * 8 + u];
    ring[i] = 0.0;
  }
  int buf = 0;
  bool pending = false;
  int pend_oy = 0;
  auto flush = [&](int b, int oy) {   // 8 windows x 384 bytes: thread t moves 16 bytes of window t / 24
    const int wdw = t / 24, seg = t % 24;
    const bool real = ox0 + wdw < pw;
    if (!LIMBS) {
      if (real)
        reinterpret_cast<uint4 *>(out + ((int64_t)oy * pw + ox0) * 192)[t] = *reinterpret_cast<const uint4 *>(&s_out[b][wdw][seg * 8]);
    } else {
      const uint4 v = real ? *reinterpret_cast<const uint4 *>(&s_out[b][wdw][seg * 8]) : make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4 *>(limbs + ((int64_t)oy * pwp + ox0) * 384)[t] = v;    // padding windows: zero rows
      if (t < SF_OXB) {   // the row's norms are complete (a barrier separates them from the adds); zero the slot for its next use
        norms[(int64_t)oy * pwp + ox0 + t] = ox0 + t < pw ? s_nrm[b][t] : 0u;
        s_nrm[b][t] = 0;
      }
    }
  };
  for (int yc = oy0; yc < y_end; yc += SF_CH) {
    __syncthreads();   // the previous chunk's planes are no longer read
    // convert the chunk: RGBToYUV (utils.pas:478-490), double evaluation, single storage (what ConvertToCpnPixels leaves)
    for (int i = t; i < SF_CH * (SF_OXB + 7); i += 192) {
      const int yl = i / (SF_OXB + 7), xl = i - yl * (SF_OXB + 7);
      const int y = yc + yl, x = ox0 + xl;
      float py = 0.f, pu = 0.f, pv = 0.f;
      if (y < fh && x < fw) {
        const int32_t col = __ldg(frame + (int64_t)y * fw + x);
        rgb_to_yuv(col & 255, (col >> 8) & 255, (col >> 16) & 255, py, pu, pv);
      }
      s_pl[0][yl][xl] = (double)py; s_pl[1][yl][xl] = (double)pu; s_pl[2][yl][xl] = (double)pv;
    }
    __syncthreads();
#pragma unroll 1
    for (int g = 0; g < SF_CH; g += 8) {
#define SF_ROW(PH)                                                                                                        \
      {                                                                                                                   \
        const int y = yc + g + PH;                                                                                        \
        if (y < y_end) {                                                                                                  \
          const bool emit = y - oy0 >= 7;                                                                                 \
          if (pending) flush(buf ^ 1, pend_oy);                                                                           \
          sf_step<PH, LIMBS>(&s_pl[cpn][g + PH][oxl], cu, sc, dst, ring, C, emit, &s_out[buf][oxl][0], &s_nrm[buf][oxl]); \
          __syncthreads();                                                                                                \
          pending = emit; pend_oy = y - 7;                                                                                \
          if (emit) buf ^= 1;                                                                                             \
        }                                                                                                                 \
      }
      SF_ROW(0) SF_ROW(1) SF_ROW(2) SF_ROW(3) SF_ROW(4) SF_ROW(5) SF_ROW(6) SF_ROW(7)
#undef SF_ROW
    }
  }
  if (pending) flush(buf ^ 1, pend_oy);
}

// Register caps, measured: the register file is four 16 K partitions, so the int16 variant at 104 registers fits only 4 warps per
// partition and a third 6-warp block does not (ncu: launch__occupancy_limit_registers = 2); capped at 96 three blocks are
// resident (0.194 -> 0.182 ms per 720p frame).  The limb variant got slower under the same cap (80 registers, 0.248 -> 0.276 ms)
// and keeps its natural 109.
__global__ void __maxnreg__(96)
features_sliding_fast_kernel(const int32_t *__restrict__ frame, int fw, int fh, int16_t *__restrict__ out) {
  sf_body<false>(frame, fw, fh, out, nullptr, nullptr, 0);
}
__global__ void __launch_bounds__(192, 2)
features_sliding_limbs_kernel(const int32_t *__restrict__ frame, int fw, int fh, uint8_t *__restrict__ limbs, uint32_t *__restrict__ norms,
                              int pwp) {
  sf_body<true>(frame, fw, fh, nullptr, limbs, norms, pwp);
}

static int features_fast_init(cudaStream_t st) {
  static bool done[TM_MAX_DEVICES] = {};
  if (!first_use_on_device(done)) return TM_OK;
  double hc[8][8], hs[192];
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < 8; ++k)
    for (int i = 0; i < 8; ++i) hc[k][i] = cos((i + 0.5) * k * PI / 8.0);
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 8; ++v)
      for (int u = 0; u < 8; ++u) {
        const float rf = (v == 0 && u == 0) ? 0.5f : ((v == 0 || u == 0) ? (float)sqrt(0.5) : 1.0f);   // cDCTUVRatio (utils.pas:100-109)
        double w;
        if (cudaMemcpyFromSymbol(&w, c_weights, 8, (size_t)(c * 64 + v * 8 + u) * 8) != cudaSuccess) return TM_ERR_CUDA;
        hs[c * 64 + v * 8 + u] = (double)rf * w;
      }
  if (cudaMemcpyToSymbolAsync(c_cos8, hc, sizeof(hc), 0, cudaMemcpyHostToDevice, st) != cudaSuccess) return TM_ERR_CUDA;
  if (cudaMemcpyToSymbolAsync(c_scale, hs, sizeof(hs), 0, cudaMemcpyHostToDevice, st) != cudaSuccess) return TM_ERR_CUDA;
  return cudaStreamSynchronize(st) == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

static int g_feature_mode = 0;   // 0: bit-exact (DCTInner_asm order), 1: fast separable f64 for the sliding-window features
void set_feature_mode(int mode) { g_feature_mode = mode ? 1 : 0; }
int get_feature_mode() { return g_feature_mode; }

static int grid_for(int64_t n, int per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t g = (int64_t)sms * per_sm;
  return (int)(n < g ? n : g);
}

int launch_features_rgb(const int32_t *rgb, int64_t n, int16_t *out, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  int rc = features_init(st);
  if (rc) return rc;
  ProfScope prof("features_rgb", st);
  features_i16_kernel<0><<<grid_for((n + FEAT_NP - 1) / FEAT_NP, 3), 192, 0, st>>>(rgb, nullptr, 0, nullptr, nullptr, 0, n, g_lutT_f32, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_features_rgb_mirrored(const int32_t *rgb, const uint8_t *flags, int64_t n, int16_t *out, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  int rc = features_init(st);
  if (rc) return rc;
  features_i16_kernel<3><<<grid_for((n + FEAT_NP - 1) / FEAT_NP, 3), 192, 0, st>>>(rgb, flags, 0, nullptr, nullptr, 0, n, g_lutT_f32, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// frame [h][w] packed RGB -> features of the 8x8 window at every pixel offset, [(h-7)][(w-7)][192]
int launch_features_sliding(const int32_t *frame, int w, int h, int16_t *out, cudaStream_t st) {
  if (w < 8 || h < 8) return TM_ERR_ARG;
  int rc = features_init(st);
  if (rc) return rc;
  const int64_t n = (int64_t)(w - 7) * (h - 7);
  ProfScope prof("features_sliding", st);
  if (g_feature_mode == 1) {
    rc = features_fast_init(st);
    if (rc) return rc;
    const dim3 grid((unsigned)((w - 7 + SF_OXB - 1) / SF_OXB), (unsigned)((h - 7 + SF_SEG - 1) / SF_SEG));
    features_sliding_fast_kernel<<<grid, 192, 0, st>>>(frame, w, h, out);
    note_launch();
    return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
  }
  features_i16_kernel<4><<<grid_for((n + FEAT_NP - 1) / FEAT_NP, 3), 192, 0, st>>>(frame, nullptr, 0, nullptr, nullptr, 0, n, g_lutT_f32, out, w,
                                                                             w - 7);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

// fast mode only: the sliding-window features of `frame` straight into the motion search's candidate operands (limb rows at
// padded positions [h - 7][pwp][384] + norms [h - 7][pwp]); see features_sliding_limbs_kernel
int launch_features_sliding_limbs(const int32_t *frame, int w, int h, uint8_t *limbs, uint32_t *norms, int pwp, cudaStream_t st) {
  if (w < 8 || h < 8 || pwp < w - 7 || (pwp % SF_OXB) != 0) return TM_ERR_ARG;
  int rc = features_init(st);
  if (rc) return rc;
  rc = features_fast_init(st);
  if (rc) return rc;
  ProfScope prof("features_sliding", st);
  const dim3 grid((unsigned)(pwp / SF_OXB), (unsigned)((h - 7 + SF_SEG - 1) / SF_SEG));
  features_sliding_limbs_kernel<<<grid, 192, 0, st>>>(frame, w, h, limbs, norms, pwp);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_features_pal(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes, int pal_size, int64_t n,
                        int16_t *out, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  int rc = features_init(st);
  if (rc) return rc;
  features_i16_kernel<1><<<grid_for((n + FEAT_NP - 1) / FEAT_NP, 3), 192, 0, st>>>(nullptr, pal_idx, 0, tile_pal, palettes, pal_size, n,
                                                                             g_lutT_f32, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_features_allpairs(const uint8_t *pal_idx, int64_t n_tiles, const int32_t *palettes, int pal_size, int n_pal,
                             int16_t *out, cudaStream_t st) {
  const int64_t n_pairs = n_tiles * n_pal;
  if (n_pairs <= 0) return TM_OK;
  int rc = features_init(st);
  if (rc) return rc;
  features_i16_kernel<2><<<grid_for((n_pairs + FEAT_NP - 1) / FEAT_NP, 3), 192, 0, st>>>(nullptr, pal_idx, n_pal, nullptr, palettes, pal_size,
                                                                                   n_pairs, g_lutT_f32, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_features_f64(const int32_t *rgb, int64_t n, int mode, int use_lab, double *out, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  int rc = features_init(st);
  if (rc) return rc;
  if (mode == 2) {   // pvsWavelets
    features_wavelet_f64_kernel<<<grid_for(n, 2), 192, 0, st>>>(rgb, n, use_lab, out);
    note_launch();
    return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
  }
  const int special = (mode == 3 || mode == 4), weighted = (mode == 1 || mode == 4);
  static bool attr[TM_MAX_DEVICES] = {};
  if (first_use_on_device(attr)) {
    if (cudaFuncSetAttribute(features_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int)sizeof(double)) != cudaSuccess)
      return TM_ERR_CUDA;
  }
  features_f64_kernel<<<grid_for(n, 2), 192, 4096 * sizeof(double), st>>>(rgb, n, weighted, use_lab, g_lutT_f64 + special * 4096, out);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

int launch_mirror_canonicalise(int32_t *rgb, int64_t n, uint8_t *flags, cudaStream_t st) {
  if (n <= 0) return TM_OK;
  mirror_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rgb, n, flags);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? TM_OK : TM_ERR_CUDA;
}

}  // namespace tmg
