/*
 * tm_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see tm_oracle.h).
 *
 * Plain-C restatement of the reference's hot path.  Build with
 *   gcc -O2 -mssse3 -ffp-contract=off -fopenmp -fPIC -shared
 * -ffp-contract=off matters: the reference's SSE code has no FMA and feature rounding
 * depends on the exact f32/f64 summation order (utils.pas:874-1035).
 */
#include "tm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <emmintrin.h>
#include <tmmintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ tables */

/* utils.pas:47-56 */
static const uint8_t kDitheringMap[64] = {
   0, 48, 12, 60,  3, 51, 15, 63,
  32, 16, 44, 28, 35, 19, 47, 31,
   8, 56,  4, 52, 11, 59,  7, 55,
  40, 24, 36, 20, 43, 27, 39, 23,
   2, 50, 14, 62,  1, 49, 13, 61,
  34, 18, 46, 30, 33, 17, 45, 29,
  10, 58,  6, 54,  9, 57,  5, 53,
  42, 26, 38, 22, 41, 25, 37, 21};

/* utils.pas:59-68 */
static const uint8_t kDCTSnake[64] = {
   0,  1,  5,  6, 14, 15, 27, 28,
   2,  4,  7, 13, 16, 26, 29, 42,
   3,  8, 12, 17, 25, 30, 41, 43,
   9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54,
  20, 22, 33, 38, 46, 51, 55, 60,
  21, 34, 37, 47, 50, 56, 59, 61,
  35, 36, 48, 49, 57, 58, 62, 63};

/* utils.pas:72-97 (daala psnrhvs CSF tables, Y / U / V) */
static const double kDCTWeights[3][8][8] = {
  {{1.6193873005, 2.2901594831, 2.08509755623, 1.48366094411, 1.00227514334, 0.678296995242, 0.466224900598, 0.3265091542},
   {2.2901594831, 1.94321815382, 2.04793073064, 1.68731108984, 1.2305666963, 0.868920337363, 0.61280991668, 0.436405793551},
   {2.08509755623, 2.04793073064, 1.34329019223, 1.09205635862, 0.875748795257, 0.670882927016, 0.501731932449, 0.372504254596},
   {1.48366094411, 1.68731108984, 1.09205635862, 0.772819797575, 0.605636379554, 0.48309405692, 0.380429446972, 0.295774038565},
   {1.00227514334, 1.2305666963, 0.875748795257, 0.605636379554, 0.448996256676, 0.352889268808, 0.283006984131, 0.226951348204},
   {0.678296995242, 0.868920337363, 0.670882927016, 0.48309405692, 0.352889268808, 0.27032073436, 0.215017739696, 0.17408067321},
   {0.466224900598, 0.61280991668, 0.501731932449, 0.380429446972, 0.283006984131, 0.215017739696, 0.168869545842, 0.136153931001},
   {0.3265091542, 0.436405793551, 0.372504254596, 0.295774038565, 0.226951348204, 0.17408067321, 0.136153931001, 0.109083846276}},
  {{1.91113096927, 2.46074210438, 1.18284184739, 1.14982565193, 1.05017074788, 0.898018824055, 0.74725392039, 0.615105596242},
   {2.46074210438, 1.58529308355, 1.21363250036, 1.38190029285, 1.33100189972, 1.17428548929, 0.996404342439, 0.830890433625},
   {1.18284184739, 1.21363250036, 0.978712413627, 1.02624506078, 1.03145147362, 0.960060382087, 0.849823426169, 0.731221236837},
   {1.14982565193, 1.38190029285, 1.02624506078, 0.861317501629, 0.801821139099, 0.751437590932, 0.685398513368, 0.608694761374},
   {1.05017074788, 1.33100189972, 1.03145147362, 0.801821139099, 0.676555426187, 0.605503172737, 0.55002013668, 0.495804539034},
   {0.898018824055, 1.17428548929, 0.960060382087, 0.751437590932, 0.605503172737, 0.514674450957, 0.454353482512, 0.407050308965},
   {0.74725392039, 0.996404342439, 0.849823426169, 0.685398513368, 0.55002013668, 0.454353482512, 0.389234902883, 0.342353999733},
   {0.615105596242, 0.830890433625, 0.731221236837, 0.608694761374, 0.495804539034, 0.407050308965, 0.342353999733, 0.295530605237}},
  {{2.03871978502, 2.62502345193, 1.26180942886, 1.11019789803, 1.01397751469, 0.867069376285, 0.721500455585, 0.593906509971},
   {2.62502345193, 1.69112867013, 1.17180569821, 1.3342742857, 1.28513006198, 1.13381474809, 0.962064122248, 0.802254508198},
   {1.26180942886, 1.17180569821, 0.944981930573, 0.990876405848, 0.995903384143, 0.926972725286, 0.820534991409, 0.706020324706},
   {1.11019789803, 1.3342742857, 0.990876405848, 0.831632933426, 0.77418706195, 0.725539939514, 0.661776842059, 0.587716619023},
   {1.01397751469, 1.28513006198, 0.995903384143, 0.77418706195, 0.653238524286, 0.584635025748, 0.531064164893, 0.478717061273},
   {0.867069376285, 1.13381474809, 0.926972725286, 0.725539939514, 0.584635025748, 0.496936637883, 0.438694579826, 0.393021669543},
   {0.721500455585, 0.962064122248, 0.820534991409, 0.661776842059, 0.531064164893, 0.438694579826, 0.375820256136, 0.330555063063},
   {0.593906509971, 0.802254508198, 0.706020324706, 0.587716619023, 0.478717061273, 0.393021669543, 0.330555063063, 0.285345396658}}};

static float  g_lut_f32[2][4096];
static double g_lut_f64[2][4096];
static double g_inv_lut[4096];
static uint32_t g_vec_inv[1024];
static volatile int g_luts_ready = 0;

/* cDCTUVRatio is declared `array of TFloat` (single): utils.pas:100-109 */
static float uv_ratio(int v, int u) {
  if (v == 0 && u == 0) return 0.5f;
  if (v == 0 || u == 0) return (float)sqrt(0.5);
  return 1.0f;
}

/* TTilingEncoder.InitLuts, tilingencoder.pas:1683-1727 */
static void init_luts(void) {
  if (g_luts_ready) return;
#ifdef _OPENMP
#pragma omp critical(tmo_luts)
#endif
  {
    if (!g_luts_ready) {
      const double PI = 3.14159265358979323846;
      for (int i = 0; i < 1024; ++i) g_vec_inv[i] = (i >> 2) ? (uint32_t)(65536 / (i >> 2)) : 0u; /* iDiv0 */
      int i = 0;
      for (int v = 0; v < 8; ++v) for (int u = 0; u < 8; ++u) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
        double r = (double)uv_ratio(v, u);
        g_lut_f64[0][i] = cos((x + 0.5) * u * PI / 8.0) * cos((y + 0.5) * v * PI / 8.0) * r;
        g_lut_f64[1][i] = cos((x + 0.5) * u * PI / 16.0) * cos((y + 0.5) * v * PI / 16.0) * r;
        g_lut_f32[0][i] = (float)g_lut_f64[0][i];
        g_lut_f32[1][i] = (float)g_lut_f64[1][i];
        ++i;
      }
      i = 0; /* loop names follow the reference: outer pair = pixel, inner pair = frequency */
      for (int v = 0; v < 8; ++v) for (int u = 0; u < 8; ++u) for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
        g_inv_lut[i] = cos((u + 0.5) * x * PI / 8.0) * cos((v + 0.5) * y * PI / 8.0) * (double)uv_ratio(y, x) * 2 / 8.0 * 2 / 8.0;
        ++i;
      }
      __sync_synchronize();
      g_luts_ready = 1;
    }
  }
}

const uint8_t *tmo_dithering_map(void) { return kDitheringMap; }
const uint8_t *tmo_dct_snake(void) { return kDCTSnake; }
const double *tmo_dct_weights(void) { return &kDCTWeights[0][0][0]; }
const float *tmo_dct_lut_f32(int special) { init_luts(); return g_lut_f32[special ? 1 : 0]; }
const double *tmo_dct_lut_f64(int special) { init_luts(); return g_lut_f64[special ? 1 : 0]; }
const double *tmo_inv_dct_lut_f64(void) { init_luts(); return g_inv_lut; }
const uint32_t *tmo_vec_inv(void) { init_luts(); return g_vec_inv; }

int tmo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm of bench.py asks for all host cores explicitly */
void tmo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }


/* ------------------------------------------------------------------ colour */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
/* Pascal Round(): to nearest, ties to even (banker's); default C rounding mode */
static inline long pround(double x) { return lrint(x); }

int32_t tmo_to_rgb(int r, int g, int b) { return (int32_t)(((b & 255) << 16) | ((g & 255) << 8) | (r & 255)); } /* utils.pas:243 */
static inline void from_rgb(int32_t c, int *r, int *g, int *b) { *r = c & 255; *g = (c >> 8) & 255; *b = (c >> 16) & 255; }

/* utils.pas:478-490.  Constants are double (0.299 is not a single); byte*double sums in double, stored to
   single; (b - yy) mixes an integer with a single -> FPC widens to the native float (double on x86_64). */
void tmo_rgb_to_yuv(int r, int g, int b, float *y, float *u, float *v) {
  float yy = (float)(r * (299.0 / 1000.0) + g * (587.0 / 1000.0) + b * (114.0 / 1000.0));
  float uu = (float)(((double)b - (double)yy) * 0.492);
  float vv = (float)(((double)r - (double)yy) * 0.877);
  *y = yy; *u = uu; *v = vv;
}

/* utils.pas:492-509 */
int32_t tmo_yuv_to_rgb(float y, float u, float v) {
  float r = (float)((double)y + (double)v * 1.13983);
  float g = (float)((double)y - (double)u * 0.39465 - (double)v * 0.58060);
  float b = (float)((double)y + (double)u * 2.03211);
  return tmo_to_rgb(clampi((int)pround(r), 0, 255), clampi((int)pround(g), 0, 255), clampi((int)pround(b), 0, 255));
}

/* utils.pas:374-410 (D50 branch active, :391) */
void tmo_rgb_to_lab(int ir, int ig, int ib, float *ol, float *oa, float *ob) {
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
  *ol = (float)((116 * (double)y) - 16);
  *oa = (float)(500 * ((double)x - (double)y));
  *ob = (float)(200 * ((double)y - (double)z));
}

/* utils.pas:422-466 */
int32_t tmo_lab_to_rgb(float ll, float aa, float bb) {
  float y = (float)(((double)ll + 16) / 116);
  float x = (float)((double)aa / 500 + (double)y);
  float z = (float)((double)y - (double)bb / 200);
  double y3 = (double)y * y * y, x3 = (double)x * x * x, z3 = (double)z * z * z;
  y = (y3 > 0.008856) ? (float)y3 : (float)(((double)y - 16.0 / 116) / 7.787);
  x = (x3 > 0.008856) ? (float)x3 : (float)(((double)x - 16.0 / 116) / 7.787);
  z = (z3 > 0.008856) ? (float)z3 : (float)(((double)z - 16.0 / 116) / 7.787);
  x = (float)(96.6797 / 100 * (double)x);
  y = (float)(100.000 / 100 * (double)y);
  z = (float)(82.5188 / 100 * (double)z);
  float r = (float)((double)x * 0.41847 + (double)y * (-0.15866) + (double)z * (-0.082835));
  float g = (float)((double)x * (-0.091169) + (double)y * 0.25243 + (double)z * 0.015708);
  float b = (float)((double)x * 0.00092090 + (double)y * (-0.0025498) + (double)z * 0.17860);
  r = (r > 0.0031308) ? (float)(1.055 * pow((double)r, 1 / 2.4) - 0.055) : (float)(12.92 * (double)r);
  g = (g > 0.0031308) ? (float)(1.055 * pow((double)g, 1 / 2.4) - 0.055) : (float)(12.92 * (double)g);
  b = (b > 0.0031308) ? (float)(1.055 * pow((double)b, 1 / 2.4) - 0.055) : (float)(12.92 * (double)b);
  return tmo_to_rgb(clampi((int)pround((double)r * 255.0), 0, 255), clampi((int)pround((double)g * 255.0), 0, 255),
                    clampi((int)pround((double)b * 255.0), 0, 255));
}

/* Win32 MulDiv: 64-bit product, rounded to nearest, halves away from zero */
static int muldiv(int a, int b, int c) {
  if (c == 0) return -1;
  long long p = (long long)a * b;
  long long ac = c < 0 ? -(long long)c : c;
  long long ap = p < 0 ? -p : p;
  long long q = (ap + ac / 2) / ac;
  return (int)(((p < 0) != (c < 0)) ? -q : q);
}

/* utils.pas:278-325 */
void tmo_rgb_to_hsv(int32_t col, uint8_t *h, uint8_t *s, uint8_t *v) {
  int rr, gg, bb;
  from_rgb(col, &rr, &gg, &bb);
  int mx = rr; if (mx < gg) mx = gg; if (mx < bb) mx = bb;
  int mn = rr; if (mn > gg) mn = gg; if (mn > bb) mn = bb;
  int hh = 0, ss = 0, ll = mx;
  if (ll != mn) {
    int delta = ll - mn;
    ss = muldiv(delta, 255, ll);
    if (rr == ll) hh = muldiv(42, gg - bb, delta);
    else if (gg == ll) hh = muldiv(42, bb - rr, delta) + 84;
    else if (bb == ll) hh = muldiv(42, rr - gg, delta) + 168;
    hh = hh % 252; /* Pascal mod: sign follows the dividend, as C */
  }
  *h = (uint8_t)(hh & 255); *s = (uint8_t)(ss & 255); *v = (uint8_t)(ll & 255);
}

/* ------------------------------------------------------------------ features */

/* tilingencoder.pas:3049-3101 */
void tmo_convert_to_cpn(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                        int from_pal, int use_lab, int hmirror, int vmirror, float cpn[3][8][8]) {
  for (int y = 0; y < 8; ++y)
    for (int x = 0; x < 8; ++x) {
      int xx = hmirror ? 7 - x : x, yy = vmirror ? 7 - y : y;
      int32_t col = from_pal ? palette[pal_idx[yy * 8 + xx]] : rgb[yy * 8 + xx];
      int r, g, b; from_rgb(col, &r, &g, &b);
      float c0, c1, c2;
      if (use_lab) tmo_rgb_to_lab(r, g, b, &c0, &c1, &c2); else tmo_rgb_to_yuv(r, g, b, &c0, &c1, &c2);
      cpn[0][y][x] = c0; cpn[1][y][x] = c1; cpn[2][y][x] = c2;
    }
}

/* DCTInner_asm, utils.pas:874-1035: 4 steps of 16; f32 products; lanes i and i+4 (and i+8, i+12) added in f32;
   widened; two f64 lane accumulators; final haddpd. */
static double dct_inner_asm(const float *c, const float *l) {
  double acc0 = 0.0, acc1 = 0.0;
  for (int s = 0; s < 4; ++s, c += 16, l += 16) {
    float p[16];
    for (int i = 0; i < 16; ++i) p[i] = c[i] * l[i];
    float a0 = p[0] + p[4], a1 = p[1] + p[5], a2 = p[2] + p[6], a3 = p[3] + p[7];
    float b0 = p[8] + p[12], b1 = p[9] + p[13], b2 = p[10] + p[14], b3 = p[11] + p[15];
    double l0 = ((double)a0 + (double)b0) + ((double)a2 + (double)b2);
    double l1 = ((double)a1 + (double)b1) + ((double)a3 + (double)b3);
    acc0 += l0; acc1 += l1;
  }
  return acc0 + acc1;
}

/* ComputeCpnPixelsPsyVisFeatures, tilingencoder.pas:3103-3131 */
void tmo_cpn_features_i16(const float cpn[3][8][8], int mode, int16_t out[TMO_DCT]) {
  init_luts();
  int special = (mode == TMO_PVS_SPE_DCT || mode == TMO_PVS_WEIGHTED_SPE_DCT);
  int weighted = (mode == TMO_PVS_WEIGHTED_DCT || mode == TMO_PVS_WEIGHTED_SPE_DCT);
  for (int c = 0; c < 3; ++c) {
    const float *lut = g_lut_f32[special];
    for (int v = 0; v < 8; ++v)
      for (int u = 0; u < 8; ++u) {
        double z = dct_inner_asm(&cpn[c][0][0], lut);
        if (weighted) z *= kDCTWeights[c][v][u];
        out[c * 64 + kDCTSnake[v * 8 + u]] = (int16_t)pround(z);
        lut += 64;
      }
  }
}

void tmo_tile_features_i16(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                           int from_pal, int hmirror, int vmirror, int16_t out[TMO_DCT]) {
  float cpn[3][8][8];
  tmo_convert_to_cpn(rgb, pal_idx, palette, from_pal, 0, hmirror, vmirror, cpn);
  tmo_cpn_features_i16(cpn, TMO_PVS_WEIGHTED_DCT, out);
}

/* ComputeTilePsyVisFeatures, tilingencoder.pas:3133-3182 (wavelet mode not restated: out of scope) */
/* WaveletGS<Double> (tilingencoder.pas:2727-2762): normalised Haar, rows then columns of the dx x dy corner of an 8-wide buffer,
   repeated on the low-pass quadrant `depth` more times; Output may alias Data (the recursion runs in place). */
static void wavelet_gs(const double *data, double *output, int dx, int dy, int depth) {
  double tx[64], ty[64];
  memset(tx, 0, sizeof tx); memset(ty, 0, sizeof ty);
  const double factor = 1.0 / sqrt(2.0);
  for (int y = 0; y < dy; ++y) {
    const int off = y * 8;
    for (int x = 0; x < dx / 2; ++x) {
      tx[x + off] = (data[x * 2 + off] + data[(x * 2 + 1) + off]) * factor;
      tx[(x + dx / 2) + off] = (data[x * 2 + off] - data[(x * 2 + 1) + off]) * factor;
    }
  }
  for (int x = 0; x < dx; ++x)
    for (int y = 0; y < dy / 2; ++y) {
      ty[x + y * 8] = (tx[x + y * 2 * 8] + tx[x + (y * 2 + 1) * 8]) * factor;
      ty[x + (y + dy / 2) * 8] = (tx[x + y * 2 * 8] - tx[x + (y * 2 + 1) * 8]) * factor;
    }
  for (int y = 0; y < dy; ++y) memcpy(output + y * 8, ty + y * 8, sizeof(double) * (size_t)dx);
  if (depth > 0) wavelet_gs(output, output, dx / 2, dy / 2, depth - 1);
}

void tmo_tile_features_f64(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                           int mode, int from_pal, int use_lab, int hmirror, int vmirror, double out[TMO_DCT]) {
  init_luts();
  float cpn[3][8][8];
  tmo_convert_to_cpn(rgb, pal_idx, palette, from_pal, use_lab, hmirror, vmirror, cpn);
  if (mode == TMO_PVS_WAVELETS) {   /* :3151-3158: three Haar levels per plane, stored through the zig-zag table like the DCT modes */
    for (int c = 0; c < 3; ++c) {
      double cd[64], loc[64];
      for (int i = 0; i < 64; ++i) cd[i] = (double)(&cpn[c][0][0])[i];
      wavelet_gs(cd, loc, 8, 8, 2);
      for (int i = 0; i < 64; ++i) out[c * 64 + kDCTSnake[i]] = loc[i];
    }
    return;
  }
  int special = (mode == TMO_PVS_SPE_DCT || mode == TMO_PVS_WEIGHTED_SPE_DCT);
  int weighted = (mode == TMO_PVS_WEIGHTED_DCT || mode == TMO_PVS_WEIGHTED_SPE_DCT);
  for (int c = 0; c < 3; ++c) {
    double cd[64];
    for (int i = 0; i < 64; ++i) cd[i] = (double)(&cpn[c][0][0])[i];
    const double *lut = g_lut_f64[special];
    for (int v = 0; v < 8; ++v)
      for (int u = 0; u < 8; ++u) {
        double z = 0.0;
        for (int i = 0; i < 64; ++i) z += cd[i] * lut[i]; /* DCTInner<PDouble>, utils.pas:782-872 */
        if (weighted) z *= kDCTWeights[c][v][u];
        out[c * 64 + kDCTSnake[v * 8 + u]] = z;
        lut += 64;
      }
  }
}

/* ComputeInvTilePsyVisFeatures, tilingencoder.pas:3184-3255 */
void tmo_inv_tile_features_f64(const double *dct, int mode, int use_lab, int32_t rgb_out[64]) {
  init_luts();
  int weighted = (mode == TMO_PVS_WEIGHTED_DCT || mode == TMO_PVS_WEIGHTED_SPE_DCT);
  double local[3][64], cpn[3][64];
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 8; ++v)
      for (int u = 0; u < 8; ++u) {
        double d = dct[kDCTSnake[v * 8 + u] + c * 64];
        local[c][v * 8 + u] = weighted ? d / kDCTWeights[c][v][u] : d;
      }
  for (int c = 0; c < 3; ++c) {
    const double *lut = g_inv_lut;
    for (int p = 0; p < 64; ++p) {
      double z = 0.0;
      for (int i = 0; i < 64; ++i) z += local[c][i] * lut[i];
      cpn[c][p] = z;
      lut += 64;
    }
  }
  for (int p = 0; p < 64; ++p) {
    float yy = (float)cpn[0][p], uu = (float)cpn[1][p], vv = (float)cpn[2][p];
    rgb_out[p] = use_lab ? tmo_lab_to_rgb(yy, uu, vv) : tmo_yuv_to_rgb(yy, uu, vv);
  }
}

void tmo_features_from_rgb_batch(const int32_t *rgb, int64_t n, int16_t *out) {
  init_luts();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) tmo_tile_features_i16(rgb + i * 64, NULL, NULL, 0, 0, 0, out + i * TMO_DCT);
}

void tmo_features_from_pal_batch(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes,
                                 int pal_size, int64_t n, int16_t *out) {
  init_luts();
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    tmo_tile_features_i16(NULL, pal_idx + i * 64, palettes + (int64_t)tile_pal[i] * pal_size, 1, 0, 0, out + i * TMO_DCT);
}

/* ------------------------------------------------------------------ mirrors */

static int zone_sum(const int32_t *rgb, int x, int y) { /* GetTileZoneSum, tilingencoder.pas:4842-4863 (w=h=4) */
  int s = 0;
  for (int j = y; j < y + 4; ++j)
    for (int i = x; i < x + 4; ++i) {
      int r, g, b; from_rgb(rgb[j * 8 + i], &r, &g, &b);
      s += r * 299 + g * 587 + b * 114; /* ToLuma, utils.pas:262 */
    }
  return s;
}
/* tilingencoder.pas:4865-4878 */
void tmo_mirror_heuristics(const int32_t *rgb, int *hm, int *vm) {
  int q00 = zone_sum(rgb, 0, 0), q01 = zone_sum(rgb, 4, 0), q10 = zone_sum(rgb, 0, 4), q11 = zone_sum(rgb, 4, 4);
  *hm = (q00 + q10) < (q01 + q11);
  *vm = (q00 + q01) < (q10 + q11);
}
void tmo_hmirror_rgb(int32_t *p) { for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) { int32_t t = p[j*8+i]; p[j*8+i] = p[j*8+7-i]; p[j*8+7-i] = t; } }
void tmo_vmirror_rgb(int32_t *p) { for (int j = 0; j < 4; ++j) for (int i = 0; i < 8; ++i) { int32_t t = p[j*8+i]; p[j*8+i] = p[(7-j)*8+i]; p[(7-j)*8+i] = t; } }
void tmo_hmirror_pal(uint8_t *p) { for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) { uint8_t t = p[j*8+i]; p[j*8+i] = p[j*8+7-i]; p[j*8+7-i] = t; } }
void tmo_vmirror_pal(uint8_t *p) { for (int j = 0; j < 4; ++j) for (int i = 0; i < 8; ++i) { uint8_t t = p[j*8+i]; p[j*8+i] = p[(7-j)*8+i]; p[(7-j)*8+i] = t; } }

/* ------------------------------------------------------------------ distance */

/* CompareEuclideanDCTPtr, utils.pas:541-557: Cardinal accumulator, wraps mod 2^32 */
uint32_t tmo_compare_euclidean_dct(const int16_t *a, const int16_t *b) {
  uint32_t r = 0;
  for (int i = 0; i < TMO_DCT; ++i) { int32_t d = (int32_t)a[i] - (int32_t)b[i]; r += (uint32_t)d * (uint32_t)d; }
  return r;
}

/* Intended semantics of CompareEuclideanDCTPtr_asm (utils.pas:559-725): psubsw / pmaddwd / horizontal add.
   The shipped asm has a register defect; this restates what it was meant to compute and is
   the CPU baseline timed in bench.py (SSE2/SSE3 only, like the reference). */
uint32_t tmo_compare_euclidean_dct_sse(const int16_t *a, const int16_t *b) {
  __m128i acc = _mm_setzero_si128();
  for (int i = 0; i < TMO_DCT; i += 8) {
    __m128i va = _mm_loadu_si128((const __m128i *)(a + i));
    __m128i vb = _mm_loadu_si128((const __m128i *)(b + i));
    __m128i d = _mm_subs_epi16(va, vb);
    acc = _mm_add_epi32(acc, _mm_madd_epi16(d, d));
  }
  acc = _mm_hadd_epi32(acc, acc);
  acc = _mm_hadd_epi32(acc, acc);
  return (uint32_t)_mm_cvtsi128_si32(acc);
}

/* QuickTestEuclideanDCTPtr, utils.pas:755-759 */
int tmo_quick_test(const int16_t *a, const int16_t *b, uint32_t min_dist) {
  uint32_t r = 0;
  for (int i = 0; i < 8; ++i) { int32_t d = (int32_t)a[i] - (int32_t)b[i]; r += (uint32_t)(d * d); }
  return r < min_dist;
}

/* EuclideanToPSNR, utils.pas:1074-1078 (Single arithmetic) */
float tmo_euclidean_to_psnr(uint32_t d) {
  float r = (float)((double)d * (1.0 / TMO_DCT));
  double m = r > 0.5 ? (double)r : 0.5;
  return (float)(10 * log10(255.0 * 255.0 / m));
}

/* ------------------------------------------------------------------ exact k-NN */

typedef struct { uint32_t d; int32_t i; } knn_ent;
static inline int ent_less(knn_ent a, knn_ent b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }

/* Contract of ann_kdtree_short_search / _search_multi with eps = 0 (extern.pas:184-185; call sites
   tilingencoder.pas:1547,1563): the exact k nearest under the uint32 distance.  ANN's tie order is unpinned;
   the oracle orders by (distance, index). */
void tmo_knn_short(const int16_t *dict, int64_t n_dict, const int16_t *q, int64_t n_q, int k,
                   int32_t *idx, uint32_t *dist, int use_sse) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t qi = 0; qi < n_q; ++qi) {
    knn_ent heap[256]; /* max-heap on (d,i) of size <= k */
    int hn = 0;
    const int16_t *qv = q + qi * TMO_DCT;
    for (int64_t di = 0; di < n_dict; ++di) {
      const int16_t *dv = dict + di * TMO_DCT;
      knn_ent e;
      e.d = use_sse ? tmo_compare_euclidean_dct_sse(qv, dv) : tmo_compare_euclidean_dct(qv, dv);
      e.i = (int32_t)di;
      if (hn < k) {
        int c = hn++;
        heap[c] = e;
        while (c > 0) { int p = (c - 1) >> 1; if (ent_less(heap[p], heap[c])) { knn_ent t = heap[p]; heap[p] = heap[c]; heap[c] = t; c = p; } else break; }
      } else if (ent_less(e, heap[0])) {
        heap[0] = e;
        int c = 0;
        for (;;) {
          int l = 2 * c + 1, r = l + 1, m = c;
          if (l < hn && ent_less(heap[m], heap[l])) m = l;
          if (r < hn && ent_less(heap[m], heap[r])) m = r;
          if (m == c) break;
          knn_ent t = heap[m]; heap[m] = heap[c]; heap[c] = t; c = m;
        }
      }
    }
    /* heap-sort ascending */
    int n = hn;
    while (n > 1) {
      knn_ent t = heap[0]; heap[0] = heap[n - 1]; heap[n - 1] = t; --n;
      int c = 0;
      for (;;) {
        int l = 2 * c + 1, r = l + 1, m = c;
        if (l < n && ent_less(heap[m], heap[l])) m = l;
        if (r < n && ent_less(heap[m], heap[r])) m = r;
        if (m == c) break;
        knn_ent t2 = heap[m]; heap[m] = heap[c]; heap[c] = t2; c = m;
      }
    }
    for (int j = 0; j < k; ++j) {
      idx[qi * k + j] = j < hn ? heap[j].i : -1;
      dist[qi * k + j] = j < hn ? heap[j].d : 0xFFFFFFFFu;
    }
  }
}

/* ann_kdtree_search (ANN.dll, extern.pas:180; call site tilingencoder.pas:4128): exact NN of doubles, eps 0 */
void tmo_knn_double(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q,
                    int32_t *idx, double *dist) {
#pragma omp parallel for schedule(static)
  for (int64_t qi = 0; qi < n_q; ++qi) {
    double best = INFINITY; int32_t bi = -1;
    for (int64_t di = 0; di < n_dict; ++di) {
      double s = 0.0;
      for (int j = 0; j < dim; ++j) { double d = q[qi * dim + j] - dict[di * dim + j]; s += d * d; } /* utils.pas:727-734 */
      if (s < best) { best = s; bi = (int32_t)di; }
    }
    idx[qi] = bi; if (dist) dist[qi] = best;
  }
}

/* ------------------------------------------------------------------ dithering */

/* PreparePlan, tilingencoder.pas:2268-2301 */
void tmo_prepare_plan(tmo_plan *plan, const int32_t *pal, int pal_len, int y2_mixed_colors) {
  memset(plan, 0, sizeof(*plan));
  plan->y2_mixed_colors = y2_mixed_colors;
  int cnt = 0;
  for (int i = 0; i < pal_len && i < 256; ++i) {
    if (pal[i] == TMO_NULL_COLOR) continue;
    int r, g, b; from_rgb(pal[i], &r, &g, &b);
    plan->luma_pal[cnt] = r * 299 + g * 587 + b * 114;
    plan->y2[cnt][0] = r; plan->y2[cnt][1] = g; plan->y2[cnt][2] = b; plan->y2[cnt][3] = plan->luma_pal[cnt] / 1000;
    plan->remap[cnt] = (uint8_t)i;
    ++cnt;
  }
  plan->count = cnt;
}

/* ColorCompare, tilingencoder.pas:2323-2337 (Int64; div truncates toward zero like C) */
int64_t tmo_color_compare(int64_t r1, int64_t g1, int64_t b1, int64_t r2, int64_t g2, int64_t b2) {
  int64_t luma1 = r1 * 299 + g1 * 587 + b1 * 114;
  int64_t luma2 = r2 * 299 + g2 * 587 + b2 * 114;
  int64_t lumadiff = (luma1 - luma2) / 1000;
  int64_t dr = r1 - r2, dg = g1 - g2, db = b1 - b2;
  return (dr * dr) * 13 + (dg * dg) * 13 + (db * db) * 13 + ((lumadiff * lumadiff) << 5);
}

/* QuickSort, extern.pas:370-418 -- middle pivot, Hoare partition, pivot index tracked through swaps;
   non-stable, so equal-luma entries land where this exact procedure puts them. */
void tmo_quicksort_bytes_by_key(uint8_t *d, int64_t first, int64_t last, const int32_t *key) {
  if (last <= first) return;
  int64_t I, J, P;
  do {
    I = first; J = last; P = (first + last) >> 1;
    do {
      while (key[d[I]] < key[d[P]]) ++I;
      while (key[d[J]] > key[d[P]]) --J;
      if (I <= J) {
        uint8_t t = d[J]; d[J] = d[I]; d[I] = t;
        if (P == I) P = J; else if (P == J) P = I;
        ++I; --J;
      }
    } while (I <= J);
    if (first < J) tmo_quicksort_bytes_by_key(d, first, J, key);
    first = I;
  } while (I < last);
}

/* DeviseBestMixingPlanThomasKnoll, tilingencoder.pas:2565-2612 */
void tmo_mixing_plan_tk(const tmo_plan *plan, int32_t col, uint8_t list[64]) {
  int r, g, b; from_rgb(col, &r, &g, &b);
  int64_t s[3] = {r, g, b}, e[3] = {0, 0, 0}, t[3];
  for (int c = 0; c < 64; ++c) {
    t[0] = s[0] + (e[0] * 9) / 100;
    t[1] = s[1] + (e[1] * 9) / 100;
    t[2] = s[2] + (e[2] * 9) / 100;
    int64_t least = INT64_MAX;
    int chosen = plan->count ? c % plan->count : 0;
    for (int i = 0; i < plan->count; ++i) {
      int64_t pen = tmo_color_compare(t[0], t[1], t[2], plan->y2[i][0], plan->y2[i][1], plan->y2[i][2]);
      if (pen < least) { least = pen; chosen = i; }
    }
    list[c] = (uint8_t)chosen;
    e[0] += s[0] - plan->y2[chosen][0];
    e[1] += s[1] - plan->y2[chosen][1];
    e[2] += s[2] - plan->y2[chosen][2];
  }
  tmo_quicksort_bytes_by_key(list, 0, 63, plan->luma_pal);
}

/* DeviseBestMixingPlanYliluoma, tilingencoder.pas:2339-2563 -- the {$define ASM_DBMP} SSE4.1 path is what ships
   (:8, :2417-2504): all FOUR lanes of `add` are incremented each step, the 4th lane carries luma/1000, the
   running mean uses the reciprocal table FVecInv ((sum * (65536 div t)) >> 16) and everything is 32-bit lanes. */
int tmo_mixing_plan_yliluoma(const tmo_plan *plan, int32_t col, uint8_t list[TMO_DITHER_LIST_LEN]) {
  init_luts();
  int r, g, b; from_rgb(col, &r, &g, &b);
  uint32_t target[4] = {(uint32_t)r, (uint32_t)g, (uint32_t)b, (uint32_t)((r * 299 + g * 587 + b * 114) / 1000)};
  static const uint32_t w[4] = {13, 13, 13, 32};
  uint32_t so_far[4] = {0, 0, 0, 0};
  int plan_count = 0;
  while (plan_count < plan->y2_mixed_colors) {
    int max_test = plan_count == 0 ? 1 : plan_count;
    uint64_t least = ((uint64_t)1 << 63) - 1;
    int chosen = 0; int chosen_t = plan_count + 1;
    for (int index = 0; index < plan->count; ++index) {
      uint32_t sum[4], add[4];
      for (int l = 0; l < 4; ++l) { sum[l] = so_far[l]; add[l] = (uint32_t)plan->y2[index][l]; }
      for (int t = plan_count + 1; t <= plan_count + max_test; ++t) {
        uint32_t inv = g_vec_inv[t * 4];
        uint32_t pen = 0;
        for (int l = 0; l < 4; ++l) {
          sum[l] += add[l];
          add[l] += 1;
          uint32_t avg = (uint32_t)(sum[l] * inv) >> 16;
          uint32_t d = avg - target[l];
          pen += (d * d) * w[l];
        }
        if ((uint64_t)pen < least) { least = pen; chosen = index; chosen_t = t; }
      }
    }
    int amount = chosen_t - plan_count;
    if (amount > TMO_DITHER_LIST_LEN - plan_count) amount = TMO_DITHER_LIST_LEN - plan_count;
    memset(list + plan_count, chosen, (size_t)amount);
    plan_count += amount;
    for (int l = 0; l < 4; ++l) so_far[l] += (uint32_t)plan->y2[chosen][l] * (uint32_t)amount;
  }
  tmo_quicksort_bytes_by_key(list, 0, plan_count - 1, plan->luma_pal);
  return plan_count;
}

/* DitherTile, tilingencoder.pas:2688-2724 */
void tmo_dither_tile(const int32_t *rgb_in, int hmirror, int vmirror, const tmo_plan *plan, int use_tk, uint8_t out_idx[64]) {
  int32_t rgb[64];
  memcpy(rgb, rgb_in, sizeof(rgb));
  if (hmirror) tmo_hmirror_rgb(rgb);   /* back to natural orientation */
  if (vmirror) tmo_vmirror_rgb(rgb);
  for (int y = 0; y < 8; ++y)
    for (int x = 0; x < 8; ++x) {
      int map_value = kDitheringMap[((y & 7) << 3) | (x & 7)];
      if (use_tk) {
        uint8_t list[64];
        tmo_mixing_plan_tk(plan, rgb[y * 8 + x], list);
        out_idx[y * 8 + x] = plan->remap[list[map_value]];
      } else {
        uint8_t list[TMO_DITHER_LIST_LEN];
        int count = tmo_mixing_plan_yliluoma(plan, rgb[y * 8 + x], list);
        map_value = (map_value * count) >> 6;
        out_idx[y * 8 + x] = plan->remap[list[map_value]];
      }
    }
  if (hmirror) tmo_hmirror_pal(out_idx);
  if (vmirror) tmo_vmirror_pal(out_idx);
}

/* TTilingEncoder.Dither, tilingencoder.pas:1873-1907, generalised to (tile, palette) pair lists:
   pair p dithers tile pair_tile[p] (or p when NULL) against palette tile_pal[p]. mirror_flags bit0=H bit1=V. */
void tmo_dither_batch(const int32_t *rgb, const uint8_t *mirror_flags, const int32_t *tile_pal, int64_t n_pairs,
                      const int32_t *pair_tile, const int32_t *palettes, int pal_size, int n_pal,
                      int use_tk, int y2_mixed_colors, uint8_t *out_idx) {
  init_luts();
  tmo_plan *plans = (tmo_plan *)malloc(sizeof(tmo_plan) * (size_t)n_pal);
  for (int p = 0; p < n_pal; ++p) tmo_prepare_plan(&plans[p], palettes + (int64_t)p * pal_size, pal_size, y2_mixed_colors);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < n_pairs; ++i) {
    int64_t t = pair_tile ? pair_tile[i] : i;
    int f = mirror_flags ? mirror_flags[t] : 0;
    tmo_dither_tile(rgb + t * 64, f & 1, (f >> 1) & 1, &plans[tile_pal[i]], use_tk, out_idx + i * 64);
  }
  free(plans);
}

/* ------------------------------------------------------------------ k-means */
#define TMO_KM_BLOCK 128

static inline uint64_t xorshift64s(uint64_t *s) {
  uint64_t x = *s; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; *s = x; return x * 0x2545F4914F6CDD1DULL;
}

/* k-means++ (Arthur & Vassilvitskii) with an explicit generator.  yakmo's initType=1 draw sequence for
   initSeed=0 (tilingencoder.pas:4198,4492) is parity unpinned: no source, no version. */
void tmo_kmeanspp_init(const double *x, int64_t n, int dim, int k, uint64_t seed, double *cent) {
  uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ULL;
  double *d2 = (double *)malloc(sizeof(double) * (size_t)n);
  int64_t first = (int64_t)(xorshift64s(&st) % (uint64_t)n);
  memcpy(cent, x + first * dim, sizeof(double) * (size_t)dim);
  for (int64_t i = 0; i < n; ++i) d2[i] = INFINITY;
  for (int c = 1; c < k; ++c) {
    const double *last = cent + (int64_t)(c - 1) * dim;
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double s = 0.0;
      for (int j = 0; j < dim; ++j) { double d = x[i * dim + j] - last[j]; s += d * d; }
      if (s < d2[i]) d2[i] = s;
      total += d2[i];
    }
    double u = (double)(xorshift64s(&st) >> 11) * (1.0 / 9007199254740992.0) * total;
    int64_t pick = n - 1; double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) { acc += d2[i]; if (acc > u) { pick = i; break; } }
    memcpy(cent + (int64_t)c * dim, x + pick * dim, sizeof(double) * (size_t)dim);
  }
  free(d2);
}

/* Lloyd iterations: the fixed-point yakmo_train_on_data converges to from a given initialisation
   (extern.pas:198-203; maxIter 300 = cYakmoMaxIterations, utils.pas:17).  Assignment = first minimum in centroid
   order; update = arithmetic mean in f64, points accumulated in index order. */
int tmo_kmeans_lloyd(const double *x, int64_t n, int dim, int k, int max_iter, double *cent,
                     int32_t *labels, double *inertia, int nan_empty) {
  double *sums = (double *)malloc(sizeof(double) * (size_t)k * dim);
  double *blk = (double *)malloc(sizeof(double) * (size_t)k * dim);
  int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
  for (int64_t i = 0; i < n; ++i) labels[i] = -1;
  int it = 0;
  double total = 0.0;
  for (;;) {
    int64_t changed = 0; total = 0.0;
#pragma omp parallel for schedule(static) reduction(+:changed, total)
    for (int64_t i = 0; i < n; ++i) {
      double best = INFINITY; int32_t bi = labels[i] >= 0 ? labels[i] : 0;
      for (int c = 0; c < k; ++c) {
        double s = 0.0;
        const double *cv = cent + (int64_t)c * dim, *xv = x + i * dim;
        for (int j = 0; j < dim; ++j) { double d = xv[j] - cv[j]; s += d * d; }
        if (s < best) { best = s; bi = c; }
      }
      if (bi != labels[i]) { labels[i] = bi; ++changed; }
      total += best;
    }
    if (changed == 0 || it >= max_iter) break;
    ++it;
    /* per-cluster sums in a FIXED blocked order: members in ascending point index, consecutive blocks of
       TMO_KM_BLOCK members summed sequentially from 0, block sums added sequentially from 0.  (yakmo's own summation
       order is unknowable; this one is parallel-friendly and is what the GPU update kernel reproduces bit for bit.
       For clusters of <= TMO_KM_BLOCK members it is the plain sequential sum.) */
    memset(sums, 0, sizeof(double) * (size_t)k * dim);
    memset(cnt, 0, sizeof(int64_t) * (size_t)k);
    memset(blk, 0, sizeof(double) * (size_t)k * dim);
    for (int64_t i = 0; i < n; ++i) {
      const int32_t c = labels[i];
      double *bv = blk + (int64_t)c * dim; const double *xv = x + i * dim;
      for (int j = 0; j < dim; ++j) bv[j] += xv[j];
      if (++cnt[c] % TMO_KM_BLOCK == 0) {
        double *sv = sums + (int64_t)c * dim;
        for (int j = 0; j < dim; ++j) { sv[j] += bv[j]; bv[j] = 0.0; }
      }
    }
    for (int c = 0; c < k; ++c)
      if (cnt[c] % TMO_KM_BLOCK != 0) {
        double *sv = sums + (int64_t)c * dim, *bv = blk + (int64_t)c * dim;
        for (int j = 0; j < dim; ++j) sv[j] += bv[j];
      }
    for (int c = 0; c < k; ++c) {
      if (cnt[c] > 0) for (int j = 0; j < dim; ++j) cent[(int64_t)c * dim + j] = sums[(int64_t)c * dim + j] / (double)cnt[c];
      else if (nan_empty) for (int j = 0; j < dim; ++j) cent[(int64_t)c * dim + j] = NAN;
    }
  }
  if (inertia) *inertia = total;
  free(sums); free(blk); free(cnt);
  return it;
}

/* Weighted Lloyd with bounded effort: the stand-in for BICO.dll's streaming coreset (bico_create / insert_line /
   get_results, extern.pas:218-223; call site tilingencoder.pas:4149-4172).  BICO's source and projections are absent
   (parity unpinned), so the coreset is DEFINED here and the GPU library reproduces it bit for bit: n <= k -> every point
   is its own summary; else unweighted k-means++ seeding (tmo_kmeanspp_init), at most max_iter weighted updates
   (centroid = sum w x / sum w, same blocked summation order as tmo_kmeans_lloyd, product rounded before the add), empty
   clusters dropped.  Returns the number of summary points; weights_out = summed weights. */
int64_t tmo_coreset_weighted(const double *x, const double *w, int64_t n, int dim, int64_t k, int max_iter, uint64_t seed,
                             double *cent_out, double *weights_out) {
  if (n <= k) {
    memcpy(cent_out, x, sizeof(double) * (size_t)n * dim);
    if (weights_out) memcpy(weights_out, w, sizeof(double) * (size_t)n);
    return n;
  }
  double *cent = (double *)malloc(sizeof(double) * (size_t)k * dim);
  double *sums = (double *)calloc((size_t)k * dim, sizeof(double)), *blk = (double *)calloc((size_t)k * dim, sizeof(double));
  double *ws = (double *)calloc((size_t)k, sizeof(double)), *bw = (double *)calloc((size_t)k, sizeof(double));
  int64_t *cnt = (int64_t *)calloc((size_t)k, sizeof(int64_t));
  int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  tmo_kmeanspp_init(x, n, dim, (int)k, seed, cent);
  for (int64_t i = 0; i < n; ++i) labels[i] = -1;
  int it = 0;
  for (;;) {
    int64_t changed = 0;
#pragma omp parallel for schedule(static) reduction(+:changed)
    for (int64_t i = 0; i < n; ++i) {
      double best = INFINITY; int32_t bi = labels[i] >= 0 ? labels[i] : 0;
      for (int64_t c = 0; c < k; ++c) {
        double s = 0.0;
        const double *cv = cent + c * dim, *xv = x + i * dim;
        for (int j = 0; j < dim; ++j) { double d = xv[j] - cv[j]; s += d * d; }
        if (s < best) { best = s; bi = (int32_t)c; }
      }
      if (bi != labels[i]) { labels[i] = bi; ++changed; }
    }
    if (changed == 0 || it >= max_iter) break;
    ++it;
    memset(sums, 0, sizeof(double) * (size_t)k * dim); memset(blk, 0, sizeof(double) * (size_t)k * dim);
    memset(ws, 0, sizeof(double) * (size_t)k); memset(bw, 0, sizeof(double) * (size_t)k);
    memset(cnt, 0, sizeof(int64_t) * (size_t)k);
    for (int64_t i = 0; i < n; ++i) {
      const int32_t c = labels[i];
      double *bv = blk + (int64_t)c * dim; const double *xv = x + i * dim;
      for (int j = 0; j < dim; ++j) { const double p = w[i] * xv[j]; bv[j] += p; }
      bw[c] += w[i];
      if (++cnt[c] % TMO_KM_BLOCK == 0) {
        double *sv = sums + (int64_t)c * dim;
        for (int j = 0; j < dim; ++j) { sv[j] += bv[j]; bv[j] = 0.0; }
        ws[c] += bw[c]; bw[c] = 0.0;
      }
    }
    for (int64_t c = 0; c < k; ++c) {
      if (cnt[c] % TMO_KM_BLOCK != 0) {
        double *sv = sums + c * dim, *bv = blk + c * dim;
        for (int j = 0; j < dim; ++j) sv[j] += bv[j];
        ws[c] += bw[c];
      }
      if (cnt[c] > 0) for (int j = 0; j < dim; ++j) cent[c * dim + j] = sums[c * dim + j] / ws[c];
    }
  }
  int64_t m = 0;
  for (int64_t c = 0; c < k; ++c) {
    if (!(it > 0 ? ws[c] > 0.0 : 1)) continue;
    memcpy(cent_out + m * dim, cent + c * dim, sizeof(double) * (size_t)dim);
    if (weights_out) weights_out[m] = it > 0 ? ws[c] : 0.0;
    ++m;
  }
  free(cent); free(sums); free(blk); free(ws); free(bw); free(cnt); free(labels);
  return m;
}

/* ------------------------------------------------------------------ palette colour quantisation */

typedef struct { int r, g, b; uint8_t h, s, v; int order; } cm_item;
static int cm_cmp(const void *pa, const void *pb) { /* CompareCountIndexVSH, utils.pas:741-748 */
  const cm_item *a = (const cm_item *)pa, *b = (const cm_item *)pb;
  if (a->v != b->v) return a->v < b->v ? -1 : 1;
  if (a->s != b->s) return a->s < b->s ? -1 : 1;
  if (a->h != b->h) return a->h < b->h ? -1 : 1;
  return a->order - b->order; /* TFPGList.Sort is unstable; full ties are unpinned -> keep centroid order */
}
static int px_cmp(const void *pa, const void *pb) { /* CompareDSPixel, tilingencoder.pas:1046-1056: (G, R, B) */
  int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
  int ar, ag, ab, br, bg, bb; from_rgb(a, &ar, &ag, &ab); from_rgb(b, &br, &bg, &bb);
  if (ag != bg) return ag - bg;
  if (ar != br) return ar - br;
  return ab - bb;
}

/* QuantizeUsingYakmo + DoQuantization, tilingencoder.pas:4434-4564 */
int tmo_quantize_palette(const int32_t *pixels, int64_t n, int pal_size, const double *init, uint64_t seed,
                         int32_t *palette_out) {
  for (int i = 0; i < pal_size; ++i) palette_out[i] = TMO_NULL_COLOR;
  if (n <= 0) return 0;
  int k = pal_size < n ? pal_size : (int)n;
  int32_t *sorted = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  memcpy(sorted, pixels, sizeof(int32_t) * (size_t)n);
  qsort(sorted, (size_t)n, sizeof(int32_t), px_cmp); /* rows equal under the key are identical: order-free */
  double *x = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  for (int64_t i = 0; i < n; ++i) { int r, g, b; from_rgb(sorted[i], &r, &g, &b); x[i*3] = r; x[i*3+1] = g; x[i*3+2] = b; }
  double *cent = (double *)calloc((size_t)k * 3, sizeof(double));
  if (k > 1) {
    int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    if (init) memcpy(cent, init, sizeof(double) * 3 * (size_t)k); else tmo_kmeanspp_init(x, n, 3, k, seed, cent);
    tmo_kmeans_lloyd(x, n, 3, k, 300, cent, labels, NULL, 1);
    free(labels);
  } else {
    for (int64_t i = 0; i < n; ++i) for (int j = 0; j < 3; ++j) cent[j] += x[i * 3 + j];
    for (int j = 0; j < 3; ++j) cent[j] /= (double)n;
  }
  cm_item *items = (cm_item *)malloc(sizeof(cm_item) * (size_t)k);
  for (int i = 0; i < k; ++i) {
    cm_item it; it.r = it.g = it.b = 0; it.order = i;
    if (!isnan(cent[i*3]) && !isnan(cent[i*3+1]) && !isnan(cent[i*3+2])) {
      it.r = clampi((int)pround(cent[i*3]), 0, 255);     /* Posterize(x, 255) is the identity, utils.pas:526-534 */
      it.g = clampi((int)pround(cent[i*3+1]), 0, 255);
      it.b = clampi((int)pround(cent[i*3+2]), 0, 255);
    }
    tmo_rgb_to_hsv(tmo_to_rgb(it.r, it.g, it.b), &it.h, &it.s, &it.v);
    items[i] = it;
  }
  qsort(items, (size_t)k, sizeof(cm_item), cm_cmp);
  for (int i = 0; i < k; ++i) palette_out[i] = tmo_to_rgb(items[i].r, items[i].g, items[i].b);
  free(items); free(cent); free(x); free(sorted);
  return k;
}

/* ------------------------------------------------------------------ matcher (no motion) */

static int int_cmp(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return x < y ? -1 : (x > y); }

/* TFrame.Reconstruct.DoXY, tilingencoder.pas:1464-1659, for a frame that starts its keyframe sequence
   (mpErr = High(Cardinal), :1495-1496): k-NN, then (extended) unique tiles x unique palettes re-rank (:1563-1609).
   The re-rank distance is the intended CompareEuclideanDCTPtr (utils.pas:541-557), not the defective asm. */
void tmo_match_tiles(const int16_t *q_feat, int64_t n_q, const int16_t *dict_feat, const uint8_t *dict_idx,
                     const int32_t *dict_pal, int64_t n_dict, const int32_t *palettes, int pal_size, int n_pal,
                     int k, int extended, tmo_match *out) {
  (void)n_pal;
  init_luts();
  int kk = extended ? k : 1;
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_q * kk);
  uint32_t *dist = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_q * kk);
  tmo_knn_short(dict_feat, n_dict, q_feat, n_q, kk, idx, dist, 0);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t qi = 0; qi < n_q; ++qi) {
    tmo_match m; m.tile_idx = -1; m.pal_idx = -1; m.err = 0xFFFFFFFFu;
    if (!extended) {
      int32_t t = idx[qi];
      if (t >= 0 && t < n_dict) { m.tile_idx = t; m.pal_idx = dict_pal[t]; m.err = dist[qi]; }
    } else {
      int32_t tiles[256], pals[256];
      for (int j = 0; j < kk; ++j) {
        int32_t t = idx[qi * kk + j];
        if (t >= 0 && t < n_dict) { tiles[j] = t; pals[j] = dict_pal[t]; } else { tiles[j] = -1; pals[j] = -1; }
      }
      qsort(tiles, (size_t)kk, sizeof(int32_t), int_cmp);
      qsort(pals, (size_t)kk, sizeof(int32_t), int_cmp);
      int32_t prev_t = -1;
      for (int a = 0; a < kk; ++a) {
        if (tiles[a] == prev_t) continue;
        int32_t prev_p = -1;
        for (int b = 0; b < kk; ++b) {
          if (pals[b] == prev_p) continue;
          int16_t f[TMO_DCT];
          tmo_tile_features_i16(NULL, dict_idx + (int64_t)tiles[a] * 64, palettes + (int64_t)pals[b] * pal_size, 1, 0, 0, f);
          if (tmo_quick_test(q_feat + qi * TMO_DCT, f, m.err)) {
            uint32_t e = tmo_compare_euclidean_dct(q_feat + qi * TMO_DCT, f);
            if (e < m.err) { m.err = e; m.tile_idx = tiles[a]; m.pal_idx = pals[b]; }
          }
          prev_p = pals[b];
        }
        prev_t = tiles[a];
      }
    }
    out[qi] = m;
  }
  free(idx); free(dist);
}

/* ------------------------------------------------------------------ motion search + Reconstruct */

/* TFrame.Reconstruct.DoDCTs / TFrame.PredictMotion.DoDCTs (tilingencoder.pas:1437-1462, 1157-1181): weighted-DCT YUV
   features of the 8x8 window at every pixel offset of a frame buffer, row-major [(h-7)][(w-7)][192]. */
void tmo_sliding_features(const int32_t *frame, int w, int h, int16_t *out) {
  init_luts();
  const int pw = w - 7, ph = h - 7;
#pragma omp parallel for schedule(dynamic, 1)
  for (int oy = 0; oy < ph; ++oy)
    for (int ox = 0; ox < pw; ++ox) {
      int32_t px[64];
      for (int y = 0; y < 8; ++y) memcpy(px + y * 8, frame + (int64_t)(oy + y) * w + ox, 8 * sizeof(int32_t));
      tmo_tile_features_i16(px, NULL, NULL, 0, 0, 0, out + ((int64_t)oy * pw + ox) * TMO_DCT);
    }
}

/* The motion search of TFrame.PredictMotion.DoXY / TFrame.Reconstruct.DoXY (tilingencoder.pas:1213-1244, 1496-1532):
   cur_feat [th*tw][192] = features of the frame's tiles in NATURAL orientation, dcts = sliding features of the previous
   (source or reconstructed) frame buffer.  radius_setting is MotionPredictRadius (the callers Dec() it, :1275, :1666).
   Row-major scan, Manhattan penalty added to the error (:1236, :1519), strict '<' keeps the first minimum.
   err = 0xFFFFFFFF and pred = (0,0) when no position was accepted. */
void tmo_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting,
                       int32_t *pred_x, int32_t *pred_y, uint32_t *err_out) {
  const int w = tw * 8, h = th * 8, pw = w - 7;
  const int R = radius_setting - 1;
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < tw * th; ++t) {
    const int sy = t / tw, sx = t % tw, dx = sx * 8, dy = sy * 8;
    const int16_t *cur = cur_feat + (int64_t)t * TMO_DCT;
    uint32_t best = 0xFFFFFFFFu; int bx = 0, by = 0;
    const int oymn = dy - R - 1 > 0 ? dy - R - 1 : 0, oymx = dy + R < h - 8 ? dy + R : h - 8;
    const int oxmn = dx - R - 1 > 0 ? dx - R - 1 : 0, oxmx = dx + R < w - 8 ? dx + R : w - 8;
    for (int oy = oymn; oy <= oymx; ++oy)
      for (int ox = oxmn; ox <= oxmx; ++ox) {
        const int16_t *p = dcts + ((int64_t)oy * pw + ox) * TMO_DCT;
        if (tmo_quick_test(cur, p, best)) {
          uint32_t e = tmo_compare_euclidean_dct(cur, p);
          e += (uint32_t)(abs(ox - dx) + abs(oy - dy));
          if (e < best) { best = e; bx = ox - dx; by = oy - dy; }
        }
      }
    pred_x[t] = bx; pred_y[t] = by; err_out[t] = best;
  }
}

/* TTilingEncoder.Reconstruct over ONE keyframe sequence (tilingencoder.pas:1928-1962 driver, 1430-1679 per frame).
   canon_tiles [n_frames][th*tw][64]: frame tiles as stored (mirror-canonicalised at load, :1393-1411), flags bit0 =
   HMirror, bit1 = VMirror.  Frame 0 is the sequence's StartFrame (no motion search, :1496).  Outputs per tile:
   tile_idx / pal_idx as the reference leaves them in the TTileMapItem (-1/-1 when the motion error is inside the dead
   band, :1534-1541), pred_x / pred_y, is_pred, err (the error whose PSNR the reference accumulates, :1614-1653), and
   optionally the reconstructed frames recon [n_frames][h][w].  CompareValue(knnErr, mpErr, 192) is restated on exact
   integers (the FPC overload taken for Cardinal arguments is not pinned by any reference test). */
void tmo_reconstruct_sequence(const int32_t *canon_tiles, const uint8_t *flags, int n_frames, int tw, int th,
                              const int16_t *dict_feat, const uint8_t *dict_idx, const int32_t *dict_pal, int64_t n_dict,
                              const int32_t *palettes, int pal_size, int n_pal, int radius_setting, int extended,
                              int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred,
                              uint32_t *err_out, int32_t *recon, double *psnr_sum) {
  init_luts();
  const int nt = tw * th, w = tw * 8, h = th * 8;
  const int64_t npos = (int64_t)(w - 7) * (h - 7);
  int32_t *buf[2];
  buf[0] = (int32_t *)calloc((size_t)w * h, 4);
  buf[1] = (int32_t *)calloc((size_t)w * h, 4);
  int16_t *dcts = (int16_t *)malloc((size_t)npos * TMO_DCT * 2);
  int16_t *ft = (int16_t *)malloc((size_t)nt * TMO_DCT * 2), *cur = (int16_t *)malloc((size_t)nt * TMO_DCT * 2);
  tmo_match *knn = (tmo_match *)malloc(sizeof(tmo_match) * (size_t)nt);
  int32_t *mx = (int32_t *)malloc(4 * (size_t)nt), *my = (int32_t *)malloc(4 * (size_t)nt);
  uint32_t *me = (uint32_t *)malloc(4 * (size_t)nt);
  double psum = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int32_t *tiles = canon_tiles + (int64_t)f * nt * 64;
    const uint8_t *fl = flags + (int64_t)f * nt;
    int32_t *front = buf[(f + 1) & 1], *back = buf[f & 1];
    tmo_features_from_rgb_batch(tiles, nt, ft);                               /* FTDCT, :1481-1482 */
    const int motion = f > 0 && radius_setting - 1 >= 0;
    if (motion) {
      tmo_sliding_features(back, w, h, dcts);
#pragma omp parallel for schedule(static)
      for (int t = 0; t < nt; ++t)                                             /* CurDCT, :1498-1499 */
        tmo_tile_features_i16(tiles + (int64_t)t * 64, NULL, NULL, 0, fl[t] & 1, (fl[t] >> 1) & 1, cur + (int64_t)t * TMO_DCT);
      tmo_motion_search(cur, tw, th, dcts, radius_setting, mx, my, me);
    }
    tmo_match_tiles(ft, nt, dict_feat, dict_idx, dict_pal, n_dict, palettes, pal_size, n_pal, 64, extended, knn);
    for (int t = 0; t < nt; ++t) {
      const int64_t o = (int64_t)f * nt + t;
      const int sy = t / tw, sx = t % tw, dx = sx * 8, dy = sy * 8;
      const uint32_t mp = motion ? me[t] : 0xFFFFFFFFu;
      uint32_t ke; int32_t ti, pi;
      if (mp <= TMO_DCT) { ti = -1; pi = -1; ke = 0xFFFFFFFFu; }               /* IsZero(mpErr, cTileDCTSize), :1534 */
      else { ti = knn[t].tile_idx; pi = knn[t].pal_idx; ke = knn[t].err; if (ti < 0) { pi = -1; ke = 0xFFFFFFFFu; } }
      const int knn_best = (uint64_t)ke + TMO_DCT < (uint64_t)mp;              /* CompareValue = LessThanValue, :1614 */
      tile_idx[o] = ti; pal_idx[o] = pi;
      pred_x[o] = motion ? mx[t] : 0; pred_y[o] = motion ? my[t] : 0;
      is_pred[o] = (uint8_t)!knn_best;
      const uint32_t e = knn_best ? ke : mp;
      err_out[o] = e;
      psum += (double)tmo_euclidean_to_psnr(e);
      if (knn_best) {                                                          /* draw fb (pal tile), :1623-1637 */
        const uint8_t *pp = dict_idx + (int64_t)ti * 64;
        const int32_t *pal = palettes + (int64_t)pi * pal_size;
        for (int ty = 0; ty < 8; ++ty) {
          const int tym = (fl[t] & 2) ? 7 - ty : ty;
          for (int tx = 0; tx < 8; ++tx) {
            const int txm = (fl[t] & 1) ? 7 - tx : tx;
            front[(int64_t)(dy + ty) * w + dx + tx] = pal[pp[tym * 8 + txm]];
          }
        }
      } else if (motion) {                                                     /* draw fb (motion predicted tile), :1646-1651 */
        for (int ty = 0; ty < 8; ++ty)
          memcpy(front + (int64_t)(dy + ty) * w + dx, back + (int64_t)(dy + ty + my[t]) * w + dx + mx[t], 8 * sizeof(int32_t));
      } else {
        for (int ty = 0; ty < 8; ++ty) memset(front + (int64_t)(dy + ty) * w + dx, 0, 8 * sizeof(int32_t));
      }
    }
    if (recon) memcpy(recon + (int64_t)f * w * h, front, (size_t)w * h * 4);
  }
  if (psnr_sum) *psnr_sum = psum;
  free(buf[0]); free(buf[1]); free(dcts); free(ft); free(cur); free(knn); free(mx); free(my); free(me);
}
