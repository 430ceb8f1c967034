/*
 * tm_gpu.h -- C ABI of libtm_gpu.so: the B200 (sm_100a) replacement for the native layer the TileMotion encoder
 * (gligli/tiler) calls on its data-parallel hot path.
 *
 * Two groups of entry points:
 *
 *  1. DROP-IN symbols: the exact exports the FreePascal host binds today in extern.pas:178-223 (ANN.dll,
 *     ANN_short.dll, yakmo.dll, BICO.dll).  Same argument meaning, same ownership rules, same "no error code"
 *     behaviour (a failed search returns an out-of-range index, which the host already maps to "no result",
 *     tilingencoder.pas:1549,1566).  ANN.dll and ANN_short.dll both export `ann_kdtree_*`; in one library the int16
 *     flavour takes the names the Pascal side already uses for it (`ann_kdtree_short_*`, extern.pas:182-185).
 *
 *  2. BATCHED symbols (tm_*): one call per stage / frame instead of one call per 8x8 tile -- what the host calls
 *     after hoisting the DLL call out of its per-tile loop (INTEGRATION.md shows the Pascal bindings).  Every array
 *     argument may be a HOST pointer or a DEVICE pointer (detected with cudaPointerGetAttributes): host arrays are
 *     staged through device scratch and the call returns with results on the host; with device arrays the work is
 *     only enqueued on the stream set by tm_set_stream and nothing is copied.
 *
 * All tm_* functions return 0 (TM_OK) or a TM_ERR_* code; tm_last_error() gives the text.  There is no CPU
 * fallback: without an sm_100 device every compute entry point fails with TM_ERR_NOGPU.
 *
 * Data conventions (those of the reference):
 *   pixel      int32 0x00BBGGRR                     (ToRGB, utils.pas:243-246)
 *   tile       64 pixels row-major [y][x]           (TRGBPixels / TPalPixels, tilingencoder.pas:101-104)
 *   feature    int16[192] = [Y|U|V][zig-zag(v,u)]   (TDCT, utils.pas:129-132)
 *   palette    int32[pal_size], unused = 0xFFFF00FF (cDitheringNullColor, utils.pas:45)
 *   mirror     bit0 = H, bit1 = V                   (tfHMirror_Initial / tfVMirror_Initial, tilingencoder.pas:120)
 */
#ifndef TM_GPU_H
#define TM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TM_OK 0
#define TM_ERR_ARG 1
#define TM_ERR_CUDA 2
#define TM_ERR_DRIVER 3
#define TM_ERR_NOGPU 4
#define TM_ERR_NOMEM 5

/* ---------------------------------------------------------------- runtime */
int tm_version(void);                      /* 100 * major + minor */
int tm_device_count(void);                 /* number of sm_100 devices visible */
int tm_set_device(int device);             /* device used by the calling thread's subsequent calls */
int tm_set_stream(void *cuda_stream);      /* cudaStream_t used by subsequent calls of this thread (NULL = default) */
const char *tm_last_error(void);
int64_t tm_kernel_launches(void);          /* kernels launched by this library since load (bench.py's gpu_launches) */
int tm_synchronize(void);
/* per-kernel CUDA-event timing on the launching stream (names: "knn_k1", "knn_topk", "rerank", "features_rgb") */
int tm_profile_enable(int on);
int tm_profile_read(const char *name, double *total_ms, int64_t *count);
/* Feature arithmetic of the SLIDING-WINDOW features (DoDCTs, tilingencoder.pas:1437-1462; tm_sliding_features,
   tm_predict_motion_frame, tm_reconstruct_*): 0 (default) = bit-exact, DCTInner_asm's summation order (utils.pas:874-1035);
   1 = fast: separable row/column DCT in f64, within 1 LSB of mode 0 on < 1e-3 of the coefficients (measured ~5e-6).  Tile
   features (tm_features_from_*) are always bit-exact.  Process-wide. */
int tm_set_feature_mode(int mode);
int tm_get_feature_mode(void);

/* ---------------------------------------------------------------- drop-in: ANN_short.dll (extern.pas:182-185) */
typedef struct tm_knn_short tm_knn_short;
/* rows: n pointers to int16[dim]; dim must be 192 (cTileDCTSize).  bucket/split are kd-tree knobs: ignored. The rows
   are copied (the host keeps them alive anyway, tilingencoder.pas:4600-4620). */
tm_knn_short *ann_kdtree_short_create(int16_t **rows, int n, int dim, int bucket, int split);
void ann_kdtree_short_destroy(tm_knn_short *h);
/* exact NN (eps ignored: always exact); returns index, *err = sum (a-b)^2 as uint32; -1 on failure */
int ann_kdtree_short_search(tm_knn_short *h, const int16_t *q, uint32_t eps, uint32_t *err);
/* k nearest, ascending (distance, index); slots beyond the dataset size get idx -1 / err 0xFFFFFFFF.
   err is the reference's Cardinal accumulator, i.e. the squared distance mod 2^32.  For k >= 2 the value 0xFFFFFFFF is also the
   kernel's marker of a masked column, so a row whose wrapped distance is EXACTLY 0xFFFFFFFF cannot enter a top-k list; that
   needs a true distance >= 2^32 - 1 (coefficients near the int16 limits -- real features are bounded by 13 212 per
   coefficient, distances by 2^31) and is the one documented deviation from the brute-force order. */
void ann_kdtree_short_search_multi(tm_knn_short *h, int *idxs, uint32_t *errs, int k, const int16_t *q, uint32_t eps);
/* The per-query searches above (and ann_kdtree_search) are thread-safe and MICRO-BATCHED: concurrent callers on one handle
   (the host's MTProcs pool threads, tilingencoder.pas:1547, 1563, 4128) are combined into one kernel launch per batch -- a
   leader takes every request queued while the previous batch was on the GPU.  queries / batches served so far: */
int tm_rendezvous_stats(tm_knn_short *h, int64_t *queries, int64_t *batches);

/* ---------------------------------------------------------------- drop-in: ANN.dll (extern.pas:178-180) */
typedef struct tm_knn_double tm_knn_double;
tm_knn_double *ann_kdtree_create(double **rows, int n, int dim, int bucket, int split);
void ann_kdtree_destroy(tm_knn_double *h);
int ann_kdtree_search(tm_knn_double *h, const double *q, double eps, double *err);

/* ---------------------------------------------------------------- drop-in: yakmo.dll (extern.pas:198-203) */
typedef struct tm_yakmo tm_yakmo;
tm_yakmo *yakmo_create(uint32_t k, uint32_t restarts, int max_iter, int init_type, int init_seed, int normalize, int verbose);
void yakmo_destroy(tm_yakmo *y);
void yakmo_set_num_threads(int n);                                         /* no-op: the GPU is the pool */
void yakmo_load_train_data(tm_yakmo *y, uint32_t rows, uint32_t cols, double **data);   /* copies */
void yakmo_train_on_data(tm_yakmo *y, int *point_to_cluster);
void yakmo_get_centroids(tm_yakmo *y, double **centroids);

/* ---------------------------------------------------------------- drop-in: BICO.dll (extern.pas:218-223) */
/* The CPU-era streaming coreset is replaced, not ported: points are buffered and get_results runs weighted Lloyd
   (seeded k-means++) on the full set on the GPU, returning <= coresetsize weighted centres -- the same contract
   (a weighted summary of at most `coresetsize` points). */
typedef struct tm_bico tm_bico;
tm_bico *bico_create(int64_t dim, int64_t n, int64_t k, int64_t nrandproj, int64_t coresetsize, int seed);
void bico_destroy(tm_bico *b);
void bico_set_num_threads(int n);
void bico_set_rebuild_properties(tm_bico *b, uint32_t interval, double initial, double grow);
void bico_insert_line(tm_bico *b, const double *row, double weight);
int64_t bico_get_results(tm_bico *b, double *centroids, double *weights);

/* ---------------------------------------------------------------- drop-in: dlquant_dll.dll (extern.pas:195-196; dlquant/quantizer.h:16-20) */
/* rgb888 image -> palette, planar R/G/B rows of 65536; 0 = ok, 1 = failure.  Bit-exact against dlquant/quantizer.c
   built with MSVC type widths (32-bit `ulong`).  lookup_bpc in 1..5. */
int dl1quant(uint8_t *inbuf, int width, int height, int quant_to, int lookup_bpc, uint8_t *userpal /* [3][65536] */);
int dl3quant(uint8_t *inbuf, int width, int height, int quant_to, int lookup_bpc, uint8_t *userpal /* [3][65536] */);
/* batched: n_img images concatenated (img_off[n_img+1] in pixels) -> palettes[n_img][quant_to][3] (R,G,B), counts[n_img] */
int tm_dl3quant_batch(const uint8_t *rgb888, const int64_t *img_off, int n_img, int quant_to, int lookup_bpc, uint8_t *palettes,
                      int32_t *counts);
int tm_dl1quant_batch(const uint8_t *rgb888, const int64_t *img_off, int n_img, int quant_to, int lookup_bpc, uint8_t *palettes,
                      int32_t *counts);

/* ---------------------------------------------------------------- batched: features (tilingencoder.pas:3049-3182) */
/* RGB tiles [n][64] -> int16 features [n][192] (ConvertToCpnPixels + ComputeCpnPixelsPsyVisFeatures, pvsWeightedDCT, YUV) */
int tm_features_from_rgb(const int32_t *rgb, int64_t n, int16_t *out);
/* RGB tiles read through mirror flags (ConvertToCpnPixels with AHMirror / AVMirror, :3049-3101): bit 0 H, bit 1 V */
int tm_features_from_rgb_mirrored(const int32_t *rgb, const uint8_t *flags, int64_t n, int16_t *out);
/* palette-index tiles [n][64] + tile_pal[n] + palettes[n_pal][pal_size] -> features (PrepareReconstruct.DoPsyV, :4570-4583) */
int tm_features_from_pal(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes, int pal_size, int n_pal,
                         int64_t n, int16_t *out);
/* ComputeTilePsyVisFeatures: f64, mode = TPsyVisMode ordinal (0 DCT, 1 weighted, 2 wavelets [3-level Haar], 3 special, 4 weighted special) */
int tm_features_f64(const int32_t *rgb, int64_t n, int mode, int use_lab, double *out);
/* load-time canonicalisation (TFrame.AsyncLoadFromImage, :1393-1411): flips tiles in place, writes mirror flags */
int tm_mirror_canonicalise(int32_t *rgb, int64_t n, uint8_t *flags);
/* sum (a-b)^2 mod 2^32 for n vector pairs (CompareEuclideanDCTPtr, utils.pas:541-557) */
int tm_distance_pairs(const int16_t *a, const int16_t *b, int64_t n, uint32_t *out);

/* ---------------------------------------------------------------- batched: k-NN */
/* dictionary from contiguous features [n][192] (host or device) */
int tm_knn_short_create(const int16_t *feat, int64_t n, tm_knn_short **out);
int tm_knn_short_destroy(tm_knn_short *h);
/* n_q queries [n_q][192] -> idx/dist [n_q][k], 1 <= k <= 64; sorted != 0 orders each row by (distance, index) */
int tm_knn_short_batch(tm_knn_short *h, const int16_t *q, int64_t n_q, int k, int32_t *idx, uint32_t *dist, int sorted);
/* exact NN of f64 vectors: dict [n_dict][dim], q [n_q][dim] -> idx[n_q], dist[n_q] (may be NULL) */
int tm_knn_double_batch(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q, int32_t *idx, double *dist);

/* ---------------------------------------------------------------- batched: dithering (tilingencoder.pas:1873-1907, 2268-2724) */
/* pair p dithers tile pair_tile[p] (p itself when pair_tile == NULL) against palette pair_pal[p].
   rgb [n_tiles][64] in stored (canonical) orientation, mirror_flags[n_tiles] or NULL.  out [n_pairs][64]. */
int tm_dither(const int32_t *rgb, const uint8_t *mirror_flags, int64_t n_tiles, const int32_t *pair_tile, const int32_t *pair_pal,
              int64_t n_pairs, const int32_t *palettes, int pal_size, int n_pal, int use_thomas_knoll, int y2_mixed_colors,
              uint8_t *out_idx);

/* ---------------------------------------------------------------- batched: k-means (yakmo semantics) */
/* Lloyd on f64 rows x[n][dim] from explicit initial centroids init[k][dim] (NULL: seeded k-means++).
   Stops when no label changes or after max_iter updates.  labels[n], centroids[k][dim] out; inertia/iters optional. */
int tm_kmeans_fit(const double *x, int64_t n, int dim, int k, int max_iter, const double *init, uint64_t seed, int nan_empty,
                  int32_t *labels, double *centroids, double *inertia, int *iters);
/* Lloyd on int16 rows x[n][192] (tile feature vectors) with f64 centroids: the nearest-centroid search runs on the tensor
   cores (exact int8-limb k-NN against the centroids rounded to int16, 4 candidates), the label is decided by exact f64
   distances, certified by an error bound; uncertified points fall to an exact f64 scan (count returned in *ambiguous).
   Labels and centroids equal those of tm_kmeans_fit on the same data.  init[k][192] is required. */
int tm_kmeans_fit_i16(const int16_t *x, int64_t n, int k, int max_iter, const double *init, int nan_empty, int32_t *labels,
                      double *centroids, double *inertia, int *iters, int64_t *ambiguous);
int tm_kmeans_partial_step_i16(const int16_t *x, int64_t n, int k, const double *centroids, int32_t *labels, double *partial_sums,
                               int64_t *partial_counts, int64_t *changed, double *inertia);
/* Persistent shard state for the sharded Lloyd loop (config C: 4 194 304 x 192 -> 262 144 centroids on 1/2/4/8 GPUs): the
   shard's points are split into limb rows once.  x [n][192] host (copied) or device (must outlive the handle).
   tm_kmeans_i16_step: one assignment of the shard against the replicated centroids[k][192] + the per-cluster partial sums
   [k][192] and counts [k] the caller all-reduces (NCCL) before tm_kmeans_finish_step.  With device arrays nothing is copied
   to the host: stats (int64[2], optional) accumulates {labels changed, points that needed the exact f64 scan}, inertia
   (optional) receives the shard's sum of squared distances. */
typedef struct tm_kmeans_i16 tm_kmeans_i16;
int tm_kmeans_i16_create(const int16_t *x, int64_t n, int k, tm_kmeans_i16 **out);
int tm_kmeans_i16_destroy(tm_kmeans_i16 *h);
int tm_kmeans_i16_step(tm_kmeans_i16 *h, const double *centroids, int32_t *labels, double *partial_sums, int64_t *partial_counts,
                       int64_t *stats, double *inertia);
/* one Lloyd step for a SHARD of the points (multi-GPU): assignment against replicated centroids, then per-cluster
   partial sums [k][dim] and counts [k] that the caller all-reduces (NCCL) before tm_kmeans_finish_step. */
int tm_kmeans_partial_step(const double *x, int64_t n, int dim, int k, const double *centroids, int32_t *labels,
                           double *partial_sums, int64_t *partial_counts, int64_t *changed, double *inertia);
int tm_kmeans_finish_step(const double *sums, const int64_t *counts, int k, int dim, int nan_empty, double *centroids);

/* the BICO stand-in in one call (bico_create / insert_line x n / get_results, tilingencoder.pas:4149-4172): <= k weighted
   summary points of the n weighted rows x[n][dim] (weight = UseCount).  n <= k: the rows themselves.  Else unweighted
   k-means++ seeding from `seed`, at most max_iter weighted Lloyd updates (bico_get_results uses 8), empty clusters dropped.
   centroids[k][dim] / weights[k] are HOST arrays (weights may be NULL); *count = number of summary points. */
int tm_coreset_weighted(const double *x, const double *w, int64_t n, int dim, int64_t k, int max_iter, uint64_t seed,
                        double *centroids, double *weights, int64_t *count);

/* ---------------------------------------------------------------- batched: palette colour quantisation (:4434-4564) */
/* all palettes at once: rgb tiles [n_tiles][64], tile_pal[n_tiles] in [0, n_pal) -> palettes[n_pal][pal_size]
   (k-means on the palette's pixels sorted by (G,R,B), centroids rounded, sorted by (V,S,H), padded with the null colour).
   init [n_pal][pal_size][3] explicit start or NULL (k-means++ from `seed`, same generator as the oracle). */
int tm_palquant_kmeans(const int32_t *rgb, const int32_t *tile_pal, int64_t n_tiles, int n_pal, int pal_size, const double *init,
                       uint64_t seed, int32_t *palettes, int *iters);

/* ---------------------------------------------------------------- batched: matcher (tilingencoder.pas:4566-4613, 1464-1659) */
typedef struct tm_matcher tm_matcher;
/* PrepareReconstruct: dictionary tiles as palette indices [n_dict][64] + their palette dict_pal[n_dict] + palettes.
   extended != 0 also builds the (tile x palette) feature table used by the extended-palette re-rank. */
int tm_matcher_create(const uint8_t *dict_idx, const int32_t *dict_pal, int64_t n_dict, const int32_t *palettes, int pal_size,
                      int n_pal, int extended, tm_matcher **out);
int tm_matcher_destroy(tm_matcher *m);
/* DoXY for tiles with no usable motion prediction (first frame of a keyframe sequence, or mpErr above the dead band):
   source tiles as RGB [n_q][64] (already canonicalised) -> TTileMapItem fields TileIdx / PalIdx and the error.
   A host batch of 8 or more k-NN waves (SMs x 128 tiles) is uploaded in four pieces (a short first one) on a second stream while the pieces
   already resident are matched; pinned host memory makes that overlap real, pageable memory still works. */
int tm_match_tiles_rgb(tm_matcher *m, const int32_t *rgb, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err);
/* mirror-variant search (north star "including H/V-mirror ... variants", BASELINE configs[4]; a superset of the reference, which
   matches the canonical orientation only): every source tile is also tried H-, V- and HV-mirrored (features of the mirrored
   tile in exact reference arithmetic); first strict minimum of the error over variant 0..3.  variant[i] bit 0 / bit 1 = the
   extra H / V mirror of the winning match: the tilemap's mirror bits are the tile's canonical flags XOR variant. */
int tm_match_tiles_rgb_mirrors(tm_matcher *m, const int32_t *rgb, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err,
                               uint8_t *variant);
/* same from precomputed features [n_q][192] */
int tm_match_tiles_feat(tm_matcher *m, const int16_t *feat, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err);
/* the dictionary's own features (device pointer, [n_dict][192]) for inspection/tests */
int tm_matcher_dict_features(tm_matcher *m, int16_t *out);


/* ---------------------------------------------------------------- batched: motion search + Reconstruct
   (TFrame.PredictMotion :1154-1282, TFrame.Reconstruct :1430-1679, TTilingEncoder.Reconstruct :1928-1962) */
/* DoDCTs: features of the 8x8 window at every pixel offset of a frame buffer [h][w] -> [(h-7)*(w-7)][192] */
int tm_sliding_features(const int32_t *frame, int w, int h, int16_t *out);
/* the window scan of DoXY: cur_feat [th*tw][192] (tiles in natural orientation) against the sliding features of the
   previous frame buffer; radius = MotionPredictRadius setting (window dy-radius .. dy+radius-1).  err includes the
   Manhattan penalty; first strict minimum of the row-major scan; err = 0xFFFFFFFF if nothing was accepted. */
int tm_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius, int32_t *pred_x, int32_t *pred_y,
                     uint32_t *err);
/* one frame of TTilingEncoder.PredictMotion: prev_frame [th*8][tw*8] pixels, this frame's stored (canonicalised) tiles
   [th*tw][64] + mirror flags -> motion vector and error per tile */
int tm_predict_motion_frame(const int32_t *prev_frame, const int32_t *canon_tiles, const uint8_t *flags, int tw, int th, int radius,
                            int32_t *pred_x, int32_t *pred_y, uint32_t *err);
/* TTilingEncoder.Reconstruct for ONE keyframe sequence (frame 0 = the sequence's start frame): canon_tiles
   [n_frames][th*tw][64], flags [n_frames][th*tw] (bit 0 HMirror, bit 1 VMirror).  Per tile: TileIdx / PalIdx / PredictedX /
   PredictedY / IsPredicted as TFrame.Reconstruct leaves them, err = the error whose PSNR it accumulates, psnr (optional)
   = EuclideanToPSNR(err), recon (optional) = the reconstructed frames [n_frames][th*8][tw*8].  The frame chain
   (sliding features of the previous reconstruction -> motion search -> decision -> draw) stays on the device. */
int tm_reconstruct_sequence(tm_matcher *m, const int32_t *canon_tiles, const uint8_t *flags, int n_frames, int tw, int th, int radius,
                            int k, int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred,
                            uint32_t *err, float *psnr, int32_t *recon);
/* one frame of the same (for hosts that keep their own frame loop): back = previous reconstructed frame buffer [th*8][tw*8]
   or NULL on the first frame of a keyframe sequence; front = this frame's reconstruction (written) */
int tm_reconstruct_frame(tm_matcher *m, const int32_t *canon_tiles, const uint8_t *flags, int tw, int th, int radius, int k,
                         const int32_t *back, int32_t *front, int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y,
                         uint8_t *is_pred, uint32_t *err, float *psnr);

/* ---------------------------------------------------------------- batched: Reduce (tilingencoder.pas:4014-4103, 4720-4781)
   Exact equivalence classes of n RGB tiles [n][64] (MakeTilesUnique on RGB pixels): class_id[n] in [0, *n_classes),
   equal ids <=> all 64 pixels equal.  The numbering is arbitrary (hash order); the host orders the chosen dictionary
   (ReindexTiles, :4626-4696). */
int tm_tile_classes(const int32_t *rgb, int64_t n, int32_t *class_id, int32_t *n_classes);

/* The tile-count search over the duplicate classes (STCGREval + GoldenRatioSearch, tilingencoder.pas:4014-4046, utils.pas:1044-1072).
   eff_psnr[n] = the tile's motion PSNR as a double (divided by 10.0 on the first frame of its keyframe sequence, :4028-4031).
   tm_reduce_class_min: per class the smallest eff_psnr of its members, all classes sorted ascending -> the number of distinct
   unpredicted tiles for a threshold x is the count of values <= x, so the host's golden-ratio search needs no pass over the tiles.
   tm_reduce_apply: for the chosen x, unpredicted[i] = !(eff_psnr[i] > x); use_count[c] = unpredicted members of class c (the
   dictionary tile's UseCount after MakeTilesUnique), first_member[c] = smallest unpredicted tile index of the class (n if none).
   tm_reduce_remap: TMI.TileIdx = new_of_class[class] for unpredicted tiles, -1 otherwise (TransferTiles + ReindexTiles remap). */
int tm_reduce_class_min(const int32_t *class_id, const double *eff_psnr, int64_t n, int64_t n_classes, double *sorted_min);
int tm_reduce_apply(const int32_t *class_id, const double *eff_psnr, int64_t n, int64_t n_classes, double x, int32_t *use_count,
                    int32_t *first_member, uint8_t *unpredicted);
int tm_reduce_remap(const int32_t *class_id, const uint8_t *unpredicted, const int32_t *new_of_class, int64_t n, int64_t n_classes,
                    int32_t *tile_idx);

/* mean squared error over the three colour channels of two packed-RGB buffers of n pixels */
int tm_mse_rgb(const int32_t *a, const int32_t *b, int64_t n, double *mse);

#ifdef __cplusplus
}
#endif
#endif
