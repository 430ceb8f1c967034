// tm_kernels.h -- internal launcher interface between the CUDA translation units and the C-ABI layer (abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define TM_OK 0
#define TM_ERR_ARG 1      // invalid argument
#define TM_ERR_CUDA 2     // CUDA runtime error (see tm_last_error)
#define TM_ERR_DRIVER 3   // driver entry point / tensor-map encoding failed
#define TM_ERR_NOGPU 4    // no sm_100 device
#define TM_ERR_NOMEM 5

#include <atomic>

namespace tmg {

// kernels launched by this library since load (reported as gpu_launches by bench.py)
extern std::atomic<long long> g_launches;
inline void note_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional per-kernel timing with CUDA events on the launching stream (tm_profile_enable / tm_profile_read):
// this is how bench.py measures the dominant kernel's duration live, outside any profiler.
void prof_begin(const char *name, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  ProfScope(const char *name, cudaStream_t s) : st(s) { prof_begin(name, s); }
  ~ProfScope() { prof_end(st); }
};

// ---- knn_i8.cu
int knn_rows_per_cta();
// norm_max (optional, device, zero-initialised by the caller): receives the largest squared norm, 0xFFFFFFFF if a row may exceed 32 bits
int launch_limb_split(const int16_t *in, int64_t n, uint8_t *limbs, uint32_t *norms, cudaStream_t st, uint32_t *norm_max = nullptr);
int launch_knn_i8(const uint8_t *q_limbs, const uint32_t *q_norm, int n_q, const uint8_t *d_limbs, const uint32_t *d_norm,
                  int n_dict, int k, int32_t *out_idx, uint32_t *out_dist, void *ws, int num_ctas, int sort_rows,
                  cudaStream_t st, const uint32_t *q_norm_max = nullptr, const uint32_t *d_norm_max = nullptr);

// ---- features.cu
int launch_features_rgb(const int32_t *rgb, int64_t n, int16_t *out, cudaStream_t st);
int launch_features_pal(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes, int pal_size, int64_t n,
                        int16_t *out, cudaStream_t st);
// features of every (dictionary tile, palette) pair: out [n_tiles][n_pal][192]
int launch_features_allpairs(const uint8_t *pal_idx, int64_t n_tiles, const int32_t *palettes, int pal_size, int n_pal,
                             int16_t *out, cudaStream_t st);
int launch_features_f64(const int32_t *rgb, int64_t n, int mode, int use_lab, double *out, cudaStream_t st);
int launch_features_rgb_mirrored(const int32_t *rgb, const uint8_t *flags, int64_t n, int16_t *out, cudaStream_t st);
int launch_features_sliding(const int32_t *frame, int w, int h, int16_t *out, cudaStream_t st);
int launch_features_sliding_limbs(const int32_t *frame, int w, int h, uint8_t *limbs, uint32_t *norms, int pwp, cudaStream_t st);

// ---- motion.cu
// motion search of every tile of a frame against the sliding features of the previous frame buffer
int launch_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting, int32_t *pred_x,
                         int32_t *pred_y, uint32_t *err, cudaStream_t st);
// the same search on the tensor cores (motion_tc.cu): ws of motion_tc_ws_bytes(tw, th) bytes
size_t motion_tc_ws_bytes(int tw, int th);
void motion_tc_cand_layout(void *ws, int tw, int th, uint8_t **c_limbs, uint32_t **c_norm, int *pwp_out);
int launch_motion_search_tc(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius_setting, int32_t *pred_x,
                            int32_t *pred_y, uint32_t *err, void *ws, size_t ws_bytes, int num_ctas, cudaStream_t st);
// TFrame.Reconstruct's decision + frame-buffer draw for one frame; motion arrays null on the first frame of a sequence
int launch_reconstruct_decide(const uint8_t *flags, int tw, int th, const int32_t *mp_x, const int32_t *mp_y, const uint32_t *mp_err,
                              const int32_t *knn_tile, const int32_t *knn_pal, const uint32_t *knn_err, const uint8_t *dict_idx,
                              const int32_t *palettes, int pal_size, const int32_t *back, int32_t *front, int32_t *tile_idx,
                              int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred, uint32_t *err, float *psnr,
                              cudaStream_t st);
// RGB PSNR accumulators: sum of squared channel differences between two packed-RGB buffers
int launch_sq_err_rgb(const int32_t *a, const int32_t *b, int64_t n, unsigned long long *acc, cudaStream_t st);
int launch_mirror_canonicalise(int32_t *rgb, int64_t n, uint8_t *flags, cudaStream_t st);
int features_init(cudaStream_t st);
void set_feature_mode(int mode);   // 0 bit-exact, 1 fast (sliding-window features only)
int get_feature_mode();

// Per-device state: function attributes (dynamic shared memory opt-in), lookup tables and scratch belong to ONE device, and a
// process may drive several (tm_set_device / torch.cuda.set_device before a call).  Callers hold the library lock.
constexpr int TM_MAX_DEVICES = 64;
inline int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < TM_MAX_DEVICES) ? d : 0;
}
// true exactly once per device for the flag array of one call site
inline bool first_use_on_device(bool (&done)[TM_MAX_DEVICES]) {
  const int d = cur_device();
  if (done[d]) return false;
  done[d] = true;
  return true;
}

// ---- dither.cu
int launch_dither(const int32_t *rgb, const uint8_t *mirror_flags, const int32_t *pair_tile, const int32_t *pair_pal,
                  int64_t n_pairs, const int32_t *palettes, int pal_size, int n_pal, int use_tk, int y2_mixed, uint8_t *out_idx,
                  cudaStream_t st);

// ---- match.cu
int launch_distance_pairs(const int16_t *a, const int16_t *b, int64_t n, uint32_t *out, cudaStream_t st);
int launch_match_rerank(const int16_t *q_feat, int64_t n_q, const int32_t *knn_idx, int k, const int32_t *dict_pal,
                        const uint8_t *dict_idx, int64_t n_dict, const int32_t *palettes, int pal_size, int n_pal,
                        const int16_t *pair_feat /* [n_dict][n_pal][192] */, const uint32_t *pair_norm /* [n_dict][n_pal]: |row|^2 mod 2^32 */,
                        int32_t *out_tile, int32_t *out_pal, uint32_t *out_err, cudaStream_t st);
int launch_row_norms(const int16_t *rows, int64_t n, uint32_t *norms, cudaStream_t st);
int launch_knn_f64(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q, int32_t *idx, double *dist,
                   cudaStream_t st);

int launch_mirror_combine(const int32_t *t4, const int32_t *p4, const uint32_t *e4, int64_t n, int32_t *out_tile, int32_t *out_pal,
                          uint32_t *out_err, uint8_t *out_variant, cudaStream_t st);

// ---- kmeans.cu
int launch_kmeans_assign_f64(const double *x, int64_t n, int dim, const double *cent, int k, int32_t *labels, double *dist,
                             int32_t *changed, cudaStream_t st);
// weights may be null (all 1); divide == 0 writes raw per-cluster sums instead of means (multi-GPU partial step)
int launch_kmeans_update_f64(const double *x, const double *weights, int64_t n, int dim, const int32_t *labels, int k, double *cent,
                             int64_t *counts, double *wsum_out, void *ws, size_t ws_bytes, int nan_empty, int divide,
                             cudaStream_t st);
// int16 points: tensor-core candidate search + exact f64 decision (kmeans.cu)
#define KMEANS_KC 4
int launch_kmeans_update_i16(const int16_t *x, int64_t n, const int32_t *labels, int k, double *cent, int64_t *counts_out, void *ws,
                             size_t ws_bytes, int nan_empty, int divide, cudaStream_t st);
int launch_round_centroids(const double *cent, int k, int16_t *out, cudaStream_t st);
int launch_kmeans_rerank_i16(const int16_t *x, int64_t n, const int32_t *cand, const uint32_t *cdist, const double *cent, int k,
                             int32_t *labels, double *dist, int32_t *changed, int32_t *amb_list, int32_t *amb_count, cudaStream_t st);
int launch_kmeans_assign_amb(const int16_t *x, const int32_t *amb_list, int n_amb, const double *cent, int k, int32_t *labels,
                             double *dist, int32_t *changed, double *xa, int32_t *la, double *da, cudaStream_t st);
int launch_amb_gather_limbs(const uint8_t *limbs, const uint32_t *norms, const int32_t *amb_list, int n_amb, uint8_t *out_limbs,
                            uint32_t *out_norms, cudaStream_t st);
int launch_kmeans_rerank64(const int16_t *x, const int32_t *amb_list, int n_amb, const int32_t *cand, const uint32_t *cdist,
                           const double *cent, int k, int32_t *labels, double *dist, int32_t *changed, int32_t *amb2_list,
                           int32_t *amb2_count, cudaStream_t st);
int launch_kmeanspp_f64(const double *x, int64_t n, int dim, int k, unsigned long long seed, double *d2_ws, double *cent,
                        cudaStream_t st);
int launch_kmeans_finish(const double *sums, const int64_t *counts, int k, int dim, int nan_empty, double *cent, cudaStream_t st);
int launch_match_plain(const int32_t *knn_idx, const uint32_t *knn_dist, int64_t n_q, const int32_t *dict_pal, int64_t n_dict,
                       int32_t *out_tile, int32_t *out_pal, uint32_t *out_err, cudaStream_t st);
size_t kmeans_update_ws_bytes(int64_t n, int k, int dim = 192);
int launch_kmeans_stats_add(int64_t *stats, const int32_t *counters, int64_t n_bf, cudaStream_t st);
int launch_sum_f64(const double *v, int64_t n, double *part, double *out, cudaStream_t st);
int run_palette_quantise(const int32_t *rgb, const int32_t *tile_pal, int64_t n_tiles, int n_pal, int pal_size, const double *init,
                         unsigned long long seed, int max_iter, int32_t *palettes_out, int32_t *iters_out, cudaStream_t st);

// ---- reduce.cu: exact equivalence classes of 256-byte tiles (MakeTilesUnique on RGB pixels)
size_t tile_classes_ws_bytes(int64_t n);
int run_tile_classes(const int32_t *rgb, int64_t n, int32_t *class_id, int32_t *n_classes_dev, void *ws, size_t ws_bytes, cudaStream_t st);

size_t class_min_ws_bytes(int64_t n_cls);
int run_class_min_sorted(const int32_t *cls, const double *eff, int64_t n, int64_t n_cls, double *sorted_min, void *ws, size_t ws_bytes,
                         cudaStream_t st);
int run_reduce_apply(const int32_t *cls, const double *eff, int64_t n, int64_t n_cls, double x, int32_t *use, int32_t *rep, uint8_t *unpred,
                     cudaStream_t st);
int run_reduce_remap(const int32_t *cls, const uint8_t *unpred, const int32_t *new_of_cls, int64_t n, int32_t *tile_idx, cudaStream_t st);

// ---- dlquant.cu
int run_dl3quant(const uint8_t *rgb, const int64_t *img_off, int n_img, int64_t max_pixels, int quant_to, int bpc, uint8_t *pal_out,
                 int32_t *count_out, cudaStream_t st);
int run_dl1quant(const uint8_t *rgb, const int64_t *img_off, int n_img, int64_t max_pixels, int quant_to, int bpc, uint8_t *pal_out,
                 int32_t *count_out, cudaStream_t st);

}  // namespace tmg
