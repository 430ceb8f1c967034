/*
 * tm_oracle.h -- CPU ORACLE for the TileMotion encoder's data-parallel core.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of what the reference
 * (gligli/tiler, FreePascal + SSE asm + ANN/yakmo/BICO/dlquant DLLs) computes on the
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call it.  The product (libtm_gpu.so) never does.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference checkout).  What pins it: the reference's own Test properties
 * (tilingencoder.pas:3847-3902), the literal tables of utils.pas:47-109, and
 * dlquant built from the reference's C (oracle/_ref).  ANN / yakmo / BICO ship as
 * source-less PE DLLs with no version pin: their exact tie order, k-means++ draw
 * sequence and coreset are PARITY UNPINNED; the oracle implements
 * the mathematical definition (exact k-NN, Lloyd from an explicit initialisation).
 */
#ifndef TM_ORACLE_H
#define TM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMO_TILE_W 8
#define TMO_TILE_PX 64
#define TMO_CPNS 3
#define TMO_DCT 192                 /* cTileDCTSize, utils.pas:40 */
#define TMO_NULL_COLOR ((int32_t)0xffff00ff) /* cDitheringNullColor, utils.pas:45 */
#define TMO_DITHER_LIST_LEN 256     /* cDitheringListLen, utils.pas:46 */

/* TPsyVisMode, tilingencoder.pas:21 */
enum { TMO_PVS_DCT = 0, TMO_PVS_WEIGHTED_DCT = 1, TMO_PVS_WAVELETS = 2, TMO_PVS_SPE_DCT = 3, TMO_PVS_WEIGHTED_SPE_DCT = 4 };

/* ---- tables (utils.pas:47-109) ---- */
const uint8_t *tmo_dithering_map(void);   /* 64 */
const uint8_t *tmo_dct_snake(void);       /* 64 */
const double  *tmo_dct_weights(void);     /* [3][8][8] */
const float   *tmo_dct_lut_f32(int special);  /* [v][u][y][x] 4096, InitLuts tilingencoder.pas:1703-1714 */
const double  *tmo_dct_lut_f64(int special);
const double  *tmo_inv_dct_lut_f64(void);     /* tilingencoder.pas:1718-1726 */
const uint32_t *tmo_vec_inv(void);            /* 1024, tilingencoder.pas:1698-1699 */

/* ---- colour maths (utils.pas:243-509) ---- */
int32_t tmo_to_rgb(int r, int g, int b);
void tmo_rgb_to_yuv(int r, int g, int b, float *y, float *u, float *v);       /* utils.pas:478-490 */
int32_t tmo_yuv_to_rgb(float y, float u, float v);                             /* utils.pas:492-509 */
void tmo_rgb_to_lab(int r, int g, int b, float *ol, float *oa, float *ob);    /* utils.pas:374-410 */
int32_t tmo_lab_to_rgb(float l, float a, float b);                             /* utils.pas:422-466 */
void tmo_rgb_to_hsv(int32_t col, uint8_t *h, uint8_t *s, uint8_t *v);         /* utils.pas:278-325 */

/* ---- features (tilingencoder.pas:3049-3255, utils.pas:874-1035) ---- */
/* rgb: 64 packed 0x00BBGGRR pixels; pal_idx+palette used when from_pal != 0 */
void tmo_convert_to_cpn(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                        int from_pal, int use_lab, int hmirror, int vmirror, float cpn[3][8][8]);
void tmo_cpn_features_i16(const float cpn[3][8][8], int mode, int16_t out[TMO_DCT]);
void tmo_tile_features_i16(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                           int from_pal, int hmirror, int vmirror, int16_t out[TMO_DCT]);
void tmo_tile_features_f64(const int32_t *rgb, const uint8_t *pal_idx, const int32_t *palette,
                           int mode, int from_pal, int use_lab, int hmirror, int vmirror, double out[TMO_DCT]);
void tmo_inv_tile_features_f64(const double *dct, int mode, int use_lab, int32_t rgb_out[64]);
/* batched helpers (OpenMP over tiles) */
void tmo_features_from_rgb_batch(const int32_t *rgb, int64_t n, int16_t *out);
void tmo_features_from_pal_batch(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes,
                                 int pal_size, int64_t n, int16_t *out);

/* ---- mirrors (tilingencoder.pas:4842-4878, 3257-3311) ---- */
void tmo_mirror_heuristics(const int32_t *rgb, int *hmirror, int *vmirror);
void tmo_hmirror_rgb(int32_t *rgb);
void tmo_vmirror_rgb(int32_t *rgb);
void tmo_hmirror_pal(uint8_t *p);
void tmo_vmirror_pal(uint8_t *p);

/* ---- distance (utils.pas:541-557, 755-759, 1074-1078) ---- */
uint32_t tmo_compare_euclidean_dct(const int16_t *a, const int16_t *b);
uint32_t tmo_compare_euclidean_dct_sse(const int16_t *a, const int16_t *b);   /* intended semantics of :559-725 */
int tmo_quick_test(const int16_t *a, const int16_t *b, uint32_t min_dist);
float tmo_euclidean_to_psnr(uint32_t d);

/* ---- exact k-NN (the ANN_short.dll contract, extern.pas:182-185; ANN.dll :178-180) ---- */
/* results ordered by (distance, index) ascending; slots beyond n_dict are idx=-1, dist=0xFFFFFFFF */
void tmo_knn_short(const int16_t *dict, int64_t n_dict, const int16_t *q, int64_t n_q, int k,
                   int32_t *idx, uint32_t *dist, int use_sse);
void tmo_knn_double(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q,
                    int32_t *idx, double *dist);

/* ---- dithering (tilingencoder.pas:2268-2337, 2339-2612, 2688-2724; extern.pas:370-418) ---- */
typedef struct {
  int count;               /* non-null colours */
  int32_t luma_pal[256];   /* LumaPal */
  int32_t y2[256][4];      /* Y2Palette */
  uint8_t remap[256];      /* Remap */
  int y2_mixed_colors;
} tmo_plan;
void tmo_prepare_plan(tmo_plan *plan, const int32_t *pal, int pal_len, int y2_mixed_colors);
int64_t tmo_color_compare(int64_t r1, int64_t g1, int64_t b1, int64_t r2, int64_t g2, int64_t b2);
void tmo_mixing_plan_tk(const tmo_plan *plan, int32_t col, uint8_t list[64]);
int  tmo_mixing_plan_yliluoma(const tmo_plan *plan, int32_t col, uint8_t list[TMO_DITHER_LIST_LEN]);
/* rgb in canonical (mirrored) orientation as stored in the tile; flags = tile's initial mirrors */
void tmo_dither_tile(const int32_t *rgb, int hmirror, int vmirror, const tmo_plan *plan, int use_tk, uint8_t out_idx[64]);
void tmo_dither_batch(const int32_t *rgb, const uint8_t *mirror_flags, const int32_t *tile_pal, int64_t n_pairs,
                      const int32_t *pair_tile, const int32_t *palettes, int pal_size, int n_pal,
                      int use_tk, int y2_mixed_colors, uint8_t *out_idx);
/* the repo's own QuickSort (extern.pas:370-418) on bytes keyed by an int table */
void tmo_quicksort_bytes_by_key(uint8_t *data, int64_t first, int64_t last, const int32_t *key);

/* ---- k-means: Lloyd from an explicit initialisation (yakmo contract, extern.pas:198-203) ---- */
/* returns iterations run; stops when no label changes or max_iter reached; empty clusters keep NaN centroid
   when nan_empty != 0 (host tolerates NaN, tilingencoder.pas:4521), else keep previous centroid */
int tmo_kmeans_lloyd(const double *x, int64_t n, int dim, int k, int max_iter, double *centroids /* in: init, out */,
                     int32_t *labels, double *inertia, int nan_empty);
/* deterministic k-means++ seeding with an explicit xorshift RNG (OUR definition; yakmo's draw order is unpinned) */
void tmo_kmeanspp_init(const double *x, int64_t n, int dim, int k, uint64_t seed, double *centroids);

/* ---- coreset: the BICO.dll stand-in (extern.pas:218-223; tilingencoder.pas:4149-4172), defined by this restatement ---- */
int64_t tmo_coreset_weighted(const double *x, const double *w, int64_t n, int dim, int64_t k, int max_iter, uint64_t seed,
                             double *cent_out, double *weights_out);

/* ---- palette colour quantisation (tilingencoder.pas:4434-4564) ---- */
/* pixels: n packed 0x00BBGGRR; init centroids explicit (k rows x 3, may be NULL -> kmeans++ with seed);
   writes pal_size entries (sorted V,S,H; padded with null colour); returns colour count */
int tmo_quantize_palette(const int32_t *pixels, int64_t n, int pal_size, const double *init, uint64_t seed,
                         int32_t *palette_out);

/* ---- matcher decision without motion (tilingencoder.pas:1464-1659, first-frame-of-sequence case) ---- */
typedef struct { int32_t tile_idx, pal_idx; uint32_t err; } tmo_match;
/* extended palette usage (k=64 + unique tiles x unique palettes re-rank) */
void tmo_match_tiles(const int16_t *q_feat, int64_t n_q, const int16_t *dict_feat, const uint8_t *dict_idx,
                     const int32_t *dict_pal, int64_t n_dict, const int32_t *palettes, int pal_size, int n_pal,
                     int k, int extended, tmo_match *out);

/* ---- motion search + Reconstruct (tilingencoder.pas:1154-1282, 1430-1679, 1928-1962) ---- */
void tmo_sliding_features(const int32_t *frame, int w, int h, int16_t *out /* [(h-7)*(w-7)][192] */);
void tmo_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting,
                       int32_t *pred_x, int32_t *pred_y, uint32_t *err_out);
void tmo_reconstruct_sequence(const int32_t *canon_tiles, const uint8_t *flags, int n_frames, int tw, int th,
                              const int16_t *dict_feat, const uint8_t *dict_idx, const int32_t *dict_pal, int64_t n_dict,
                              const int32_t *palettes, int pal_size, int n_pal, int radius_setting, int extended,
                              int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred,
                              uint32_t *err_out, int32_t *recon, double *psnr_sum);

int tmo_num_threads(void);
void tmo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
