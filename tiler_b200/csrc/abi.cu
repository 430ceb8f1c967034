// abi.cu -- the C ABI of libtm_gpu.so (include/tm_gpu.h): drop-in exports for the DLLs the FreePascal host binds
// (extern.pas:178-223) and the batched tm_* entry points.  Host-side plumbing only: argument checks, host<->device
// staging, handle lifetime; every computation is a kernel from knn_i8.cu / features.cu / dither.cu / match.cu /
// kmeans.cu.  No CPU implementation of any stage lives here.
#include "../../include/tm_gpu.h"
#include "tm_kernels.h"

#include <cuda_runtime.h>
#include <mutex>
#include <condition_variable>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <algorithm>

namespace tmg {
std::atomic<long long> g_launches{0};

struct ProfRec { std::string name; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
void prof_begin(const char *name, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r;
  r.name = name;
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, st);
}
}
using namespace tmg;

// ------------------------------------------------------------------ thread state / errors
static thread_local cudaStream_t t_stream = nullptr;
static thread_local std::string t_err;
static std::recursive_mutex g_mu;   // one batched call at a time per process (the GPU serialises them anyway)

static int fail(int code, const char *what) {
  char buf[512];
  cudaError_t ce = cudaGetLastError();
  snprintf(buf, sizeof buf, "%s%s%s", what, ce != cudaSuccess ? ": " : "", ce != cudaSuccess ? cudaGetErrorString(ce) : "");
  t_err = buf;
  return code;
}
#define CU(x) do { if ((x) != cudaSuccess) return fail(TM_ERR_CUDA, #x); } while (0)
#define RC(x) do { int _rc = (x); if (_rc != TM_OK) return fail(_rc, #x); } while (0)

// The current device of the calling thread is checked (and its memory pool configured) once PER DEVICE: a process may drive
// several (tm_set_device, or the caller's own cudaSetDevice between calls).
static int g_gpu_ok[TM_MAX_DEVICES + 1];   // 0 unknown, 1 usable, -1 not an sm_100 device; [TM_MAX_DEVICES] = "no device at all"
static int require_gpu() {
  int n = 0;
  int &none = g_gpu_ok[TM_MAX_DEVICES];
  if (none == 0) none = (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) ? -1 : 1;
  if (none < 0) { cudaGetLastError(); return fail(TM_ERR_NOGPU, "no sm_100 (B200) device: libtm_gpu has no CPU fallback"); }
  const int dev = cur_device();
  if (g_gpu_ok[dev] == 0) {
    cudaDeviceProp p;
    g_gpu_ok[dev] = (cudaGetDeviceProperties(&p, dev) == cudaSuccess && p.major == 10) ? 1 : -1;
    if (g_gpu_ok[dev] > 0) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    }
  }
  if (g_gpu_ok[dev] < 0) return fail(TM_ERR_NOGPU, "no sm_100 (B200) device: libtm_gpu has no CPU fallback");
  return TM_OK;
}

static bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Staging of one call's arguments: host arrays get a stream-ordered device twin; outputs are copied back at finish().
struct Stage {
  cudaStream_t st;
  std::vector<void *> temps;
  struct Out { void *host; void *dev; size_t bytes; };
  std::vector<Out> outs;
  bool any_host = false;
  int err = TM_OK;
  explicit Stage(cudaStream_t s) : st(s) {}
  void *temp(size_t bytes) {
    void *d = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMallocAsync(&d, bytes, st) != cudaSuccess) { err = TM_ERR_NOMEM; return nullptr; }
    temps.push_back(d);
    return d;
  }
  template <class T> const T *in(const T *p, size_t count) {
    if (!p || is_device_ptr(p)) return p;
    any_host = true;
    void *d = temp(count * sizeof(T));
    if (d && cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) err = TM_ERR_CUDA;
    return (const T *)d;
  }
  template <class T> T *out(T *p, size_t count) {
    if (!p || is_device_ptr(p)) return p;
    any_host = true;
    void *d = temp(count * sizeof(T));
    outs.push_back({p, d, count * sizeof(T)});
    return (T *)d;
  }
  template <class T> T *inout(T *p, size_t count) {
    if (!p || is_device_ptr(p)) return p;
    T *d = out(p, count);
    if (d && cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) err = TM_ERR_CUDA;
    return d;
  }
  int finish(bool force_sync = false) {
    int rc = err;
    for (auto &o : outs)
      if (rc == TM_OK && cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = TM_ERR_CUDA;
    for (void *t : temps) cudaFreeAsync(t, st);
    temps.clear();
    if ((any_host || force_sync) && cudaStreamSynchronize(st) != cudaSuccess) rc = rc == TM_OK ? TM_ERR_CUDA : rc;
    return rc;
  }
  ~Stage() { for (void *t : temps) cudaFreeAsync(t, st); }
};

static int num_sms() {
  static int sms[TM_MAX_DEVICES] = {};
  const int dev = cur_device();
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

// ------------------------------------------------------------------ handles
// Micro-batching rendezvous of the per-query drop-in searches.  The host calls ann_kdtree_*_search once per 8x8
// tile from MaxThreadCount pool threads (tilingencoder.pas:1547, 1563, 4128); one kernel launch per query cannot feed a GPU.
// Requests queue on the handle; the first thread that finds no batch in flight becomes the leader, takes EVERYTHING queued
// (its own request included), runs one batched search and wakes the others.  While a batch is on the GPU the next requests
// accumulate, so T concurrent callers settle at ~T queries per launch without any timed wait, and a single caller pays nothing.
struct KnnReq { const void *q; int k; int32_t *idx; void *dist; bool done; };
struct Rendezvous {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<KnnReq *> pending;
  bool leader = false;
  long long batches = 0, queries = 0;   // statistics (tm_rendezvous_stats)
  template <class Exec> void submit(KnnReq &r, Exec exec) {
    std::unique_lock<std::mutex> lk(mu);
    pending.push_back(&r);
    while (!r.done) {
      if (!leader) {
        leader = true;
        std::vector<KnnReq *> batch;
        batch.swap(pending);
        lk.unlock();
        exec(batch);
        lk.lock();
        for (KnnReq *b : batch) b->done = true;
        ++batches; queries += (long long)batch.size();
        leader = false;
        cv.notify_all();
      } else {
        cv.wait(lk);
      }
    }
  }
};
struct tm_knn_short {
  int64_t n = 0;
  uint8_t *limbs = nullptr;    // [n][384]
  uint32_t *norms = nullptr;   // [ceil64(n)]
  uint32_t *norm_max = nullptr;   // largest exact squared norm of the rows (device scalar, saturated)
  Rendezvous rv;
};
struct tm_knn_double {
  int64_t n = 0; int dim = 0;
  double *pts = nullptr;
  Rendezvous rv;
};
struct tm_yakmo {
  uint32_t k = 0; int max_iter = 300; uint64_t seed = 0;
  uint32_t rows = 0, cols = 0;
  std::vector<double> data, cent;
};
struct tm_bico {
  int64_t dim = 0, k = 0, coreset = 0; uint64_t seed = 0;
  std::vector<double> rows, weights;
};
struct tm_matcher {
  int64_t n_dict = 0; int n_pal = 0, pal_size = 0, extended = 0;
  uint8_t *dict_idx = nullptr; int32_t *dict_pal = nullptr; int32_t *palettes = nullptr;
  int16_t *dict_feat = nullptr; int16_t *pair_feat = nullptr; uint32_t *pair_norm = nullptr;   // pair_norm[t * n_pal + p] = |pair_feat row|^2 mod 2^32
  tm_knn_short *knn = nullptr;
};


// ------------------------------------------------------------------ runtime
extern "C" int tm_version(void) { return 100; }
extern "C" int tm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) { cudaDeviceProp p; if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok; }
  return ok;
}
extern "C" int tm_set_device(int device) { CU(cudaSetDevice(device)); return TM_OK; }
extern "C" int tm_set_stream(void *s) { t_stream = (cudaStream_t)s; return TM_OK; }
extern "C" const char *tm_last_error(void) { return t_err.c_str(); }
extern "C" int64_t tm_kernel_launches(void) { return (int64_t)g_launches.load(); }
extern "C" int tm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return TM_OK;
}
// total milliseconds and launch count of the kernels recorded under `name` since the last read (synchronises)
extern "C" int tm_profile_read(const char *name, double *total_ms, int64_t *count) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0;
  int64_t n = 0;
  std::vector<ProfRec> keep;
  for (auto &r : g_prof) {
    if (name && r.name != name) { keep.push_back(r); continue; }
    float ms = 0.f;
    cudaEventSynchronize(r.e1);
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { tot += ms; ++n; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  cudaGetLastError();
  g_prof.swap(keep);
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return TM_OK;
}
extern "C" int tm_synchronize(void) { CU(cudaStreamSynchronize(t_stream)); return TM_OK; }
extern "C" int tm_set_feature_mode(int mode) {
  if (mode != 0 && mode != 1) return fail(TM_ERR_ARG, "tm_set_feature_mode: 0 (bit-exact) or 1 (fast)");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  set_feature_mode(mode);
  return TM_OK;
}
extern "C" int tm_get_feature_mode(void) { return get_feature_mode(); }

// ------------------------------------------------------------------ features
extern "C" int tm_features_from_rgb(const int32_t *rgb, int64_t n, int16_t *out) {
  RC(require_gpu());
  if (n < 0 || (n > 0 && (!rgb || !out))) return fail(TM_ERR_ARG, "tm_features_from_rgb: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n * 64);
  int16_t *d_out = s.out(out, (size_t)n * 192);
  if (s.err == TM_OK) s.err = launch_features_rgb(d_rgb, n, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_features_from_pal(const uint8_t *pal_idx, const int32_t *tile_pal, const int32_t *palettes, int pal_size, int n_pal,
                                    int64_t n, int16_t *out) {
  RC(require_gpu());
  if (n < 0 || pal_size < 1 || n_pal < 1 || (n > 0 && (!pal_idx || !tile_pal || !palettes || !out)))
    return fail(TM_ERR_ARG, "tm_features_from_pal: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const uint8_t *d_idx = s.in(pal_idx, (size_t)n * 64);
  const int32_t *d_tp = s.in(tile_pal, (size_t)n);
  const int32_t *d_pal = s.in(palettes, (size_t)n_pal * pal_size);
  int16_t *d_out = s.out(out, (size_t)n * 192);
  if (s.err == TM_OK) s.err = launch_features_pal(d_idx, d_tp, d_pal, pal_size, n, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_features_f64(const int32_t *rgb, int64_t n, int mode, int use_lab, double *out) {
  RC(require_gpu());
  if (n < 0 || mode < 0 || mode > 4 || (n > 0 && (!rgb || !out))) return fail(TM_ERR_ARG, "tm_features_f64: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n * 64);
  double *d_out = s.out(out, (size_t)n * 192);
  if (s.err == TM_OK) s.err = launch_features_f64(d_rgb, n, mode, use_lab, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_mirror_canonicalise(int32_t *rgb, int64_t n, uint8_t *flags) {
  RC(require_gpu());
  if (n < 0 || (n > 0 && (!rgb || !flags))) return fail(TM_ERR_ARG, "tm_mirror_canonicalise: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  int32_t *d_rgb = s.inout(rgb, (size_t)n * 64);
  uint8_t *d_fl = s.out(flags, (size_t)n);
  if (s.err == TM_OK) s.err = launch_mirror_canonicalise(d_rgb, n, d_fl, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_distance_pairs(const int16_t *a, const int16_t *b, int64_t n, uint32_t *out) {
  RC(require_gpu());
  if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(TM_ERR_ARG, "tm_distance_pairs: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *da = s.in(a, (size_t)n * 192), *db = s.in(b, (size_t)n * 192);
  uint32_t *d_out = s.out(out, (size_t)n);
  if (s.err == TM_OK) s.err = launch_distance_pairs(da, db, n, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

// ------------------------------------------------------------------ k-NN (int16)
static int knn_short_create_dev(const int16_t *d_feat, int64_t n, cudaStream_t st, tm_knn_short **out) {
  tm_knn_short *h = new tm_knn_short();
  h->n = n;
  const int64_t npad = (n + 63) / 64 * 64;
  // handle memory comes from the device's stream-ordered pool (release threshold = never): a long-running encoder re-creates
  // dictionaries of the same size over and over, and a raw cudaMalloc / cudaFree of hundreds of MB costs 0.1-0.3 s each time
  if (cudaMallocAsync((void **)&h->limbs, (size_t)(n > 0 ? n : 1) * 384, st) != cudaSuccess ||
      cudaMallocAsync((void **)&h->norms, (size_t)(npad > 0 ? npad : 64) * 4, st) != cudaSuccess ||
      cudaMallocAsync((void **)&h->norm_max, 4, st) != cudaSuccess) {
    cudaFreeAsync(h->limbs, st); cudaFreeAsync(h->norms, st); delete h;
    return TM_ERR_NOMEM;
  }
  cudaMemsetAsync(h->norms, 0, (size_t)(npad > 0 ? npad : 64) * 4, st);
  cudaMemsetAsync(h->norm_max, 0, 4, st);
  int rc = launch_limb_split(d_feat, n, h->limbs, h->norms, st, h->norm_max);
  if (rc != TM_OK) { cudaFreeAsync(h->limbs, st); cudaFreeAsync(h->norms, st); cudaFreeAsync(h->norm_max, st); delete h; return rc; }
  *out = h;
  return TM_OK;
}

extern "C" int tm_knn_short_create(const int16_t *feat, int64_t n, tm_knn_short **out) {
  RC(require_gpu());
  if (n < 1 || n > 0x7fffffff || !feat || !out) return fail(TM_ERR_ARG, "tm_knn_short_create: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *d_feat = s.in(feat, (size_t)n * 192);
  if (s.err == TM_OK) s.err = knn_short_create_dev(d_feat, n, s.st, out);
  RC(s.finish(true));
  return TM_OK;
}

extern "C" int tm_knn_short_destroy(tm_knn_short *h) {
  if (!h) return TM_OK;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaFreeAsync(h->limbs, t_stream); cudaFreeAsync(h->norms, t_stream); cudaFreeAsync(h->norm_max, t_stream);
  delete h;
  return TM_OK;
}

// queries already on the device
static int knn_short_batch_dev(tm_knn_short *h, const int16_t *d_q, int64_t n_q, int k, int32_t *d_idx, uint32_t *d_dist, int sorted,
                               Stage &s) {
  if (n_q == 0) return TM_OK;
  uint8_t *q_limbs = (uint8_t *)s.temp((size_t)n_q * 384);
  uint32_t *q_norm = (uint32_t *)s.temp((size_t)n_q * 4);
  uint32_t *q_nmax = (uint32_t *)s.temp(4);
  if (s.err) return s.err;
  if (cudaMemsetAsync(q_nmax, 0, 4, s.st) != cudaSuccess) return TM_ERR_CUDA;
  int rc = launch_limb_split(d_q, n_q, q_limbs, q_norm, s.st, q_nmax);
  if (rc) return rc;
  return launch_knn_i8(q_limbs, q_norm, (int)n_q, h->limbs, h->norms, (int)h->n, k, d_idx, d_dist, nullptr, num_sms(), sorted, s.st, q_nmax,
                       h->norm_max);
}

extern "C" int tm_knn_short_batch(tm_knn_short *h, const int16_t *q, int64_t n_q, int k, int32_t *idx, uint32_t *dist, int sorted) {
  RC(require_gpu());
  if (!h || n_q < 0 || n_q > 0x7fffffff || k < 1 || k > 64 || (n_q > 0 && (!q || !idx || !dist)))
    return fail(TM_ERR_ARG, "tm_knn_short_batch: bad argument (1 <= k <= 64)");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *d_q = s.in(q, (size_t)n_q * 192);
  int32_t *d_idx = s.out(idx, (size_t)n_q * k);
  uint32_t *d_dist = s.out(dist, (size_t)n_q * k);
  if (s.err == TM_OK) s.err = knn_short_batch_dev(h, d_q, n_q, k, d_idx, d_dist, sorted, s);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_knn_double_batch(const double *dict, int64_t n_dict, int dim, const double *q, int64_t n_q, int32_t *idx, double *dist) {
  RC(require_gpu());
  if (n_dict < 1 || dim < 1 || n_q < 0 || !dict || (n_q > 0 && (!q || !idx))) return fail(TM_ERR_ARG, "tm_knn_double_batch: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const double *d_dict = s.in(dict, (size_t)n_dict * dim), *d_q = s.in(q, (size_t)n_q * dim);
  int32_t *d_idx = s.out(idx, (size_t)n_q);
  double *d_dist = s.out(dist, (size_t)n_q);
  if (s.err == TM_OK) s.err = launch_knn_f64(d_dict, n_dict, dim, d_q, n_q, d_idx, d_dist, s.st);
  RC(s.finish());
  return TM_OK;
}

// ------------------------------------------------------------------ dithering
extern "C" int tm_dither(const int32_t *rgb, const uint8_t *mirror_flags, int64_t n_tiles, const int32_t *pair_tile, const int32_t *pair_pal,
                         int64_t n_pairs, const int32_t *palettes, int pal_size, int n_pal, int use_tk, int y2_mixed, uint8_t *out_idx) {
  RC(require_gpu());
  if (n_tiles < 0 || n_pairs < 0 || pal_size < 1 || pal_size > 256 || n_pal < 1 || y2_mixed < 1 || y2_mixed > 16 ||
      (n_pairs > 0 && (!rgb || !pair_pal || !palettes || !out_idx)) || (!pair_tile && n_pairs > n_tiles))
    return fail(TM_ERR_ARG, "tm_dither: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n_tiles * 64);
  const uint8_t *d_fl = s.in(mirror_flags, (size_t)n_tiles);
  const int32_t *d_pt = s.in(pair_tile, (size_t)n_pairs), *d_pp = s.in(pair_pal, (size_t)n_pairs);
  const int32_t *d_pal = s.in(palettes, (size_t)n_pal * pal_size);
  uint8_t *d_out = s.out(out_idx, (size_t)n_pairs * 64);
  if (s.err == TM_OK) s.err = launch_dither(d_rgb, d_fl, d_pt, d_pp, n_pairs, d_pal, pal_size, n_pal, use_tk, y2_mixed, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

// ------------------------------------------------------------------ k-means
static int kmeans_fit_dev(const double *d_x, const double *d_w, int64_t n, int dim, int k, int max_iter, const double *d_init, uint64_t seed,
                          int nan_empty, int32_t *d_labels, double *d_cent, double *d_wsum, double *inertia, int *iters, Stage &s) {
  int32_t *d_changed = (int32_t *)s.temp(4);
  double *d_dist = (double *)s.temp((size_t)n * 8);
  const size_t ws_bytes = kmeans_update_ws_bytes(n, k, dim);
  void *ws = s.temp(ws_bytes);
  if (s.err) return s.err;
  if (d_init) { if (cudaMemcpyAsync(d_cent, d_init, (size_t)k * dim * 8, cudaMemcpyDeviceToDevice, s.st) != cudaSuccess) return TM_ERR_CUDA; }
  else { int rc = launch_kmeanspp_f64(d_x, n, dim, k, seed, d_dist, d_cent, s.st); if (rc) return rc; }
  if (cudaMemsetAsync(d_labels, 0xFF, (size_t)n * 4, s.st) != cudaSuccess) return TM_ERR_CUDA;
  int it = 0;
  for (;;) {
    int32_t h_changed = 0;
    if (cudaMemsetAsync(d_changed, 0, 4, s.st) != cudaSuccess) return TM_ERR_CUDA;
    int rc = launch_kmeans_assign_f64(d_x, n, dim, d_cent, k, d_labels, d_dist, d_changed, s.st);
    if (rc) return rc;
    if (cudaMemcpyAsync(&h_changed, d_changed, 4, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) return TM_ERR_CUDA;
    if (cudaStreamSynchronize(s.st) != cudaSuccess) return TM_ERR_CUDA;
    if (h_changed == 0 || it >= max_iter) break;
    ++it;
    rc = launch_kmeans_update_f64(d_x, d_w, n, dim, d_labels, k, d_cent, nullptr, d_wsum, ws, ws_bytes, nan_empty, 1, s.st);
    if (rc) return rc;
  }
  if (iters) *iters = it;
  if (inertia) {
    std::vector<double> h((size_t)n);
    if (cudaMemcpyAsync(h.data(), d_dist, (size_t)n * 8, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) return TM_ERR_CUDA;
    if (cudaStreamSynchronize(s.st) != cudaSuccess) return TM_ERR_CUDA;
    double t = 0.0;
    for (double v : h) t += v;   // reporting scalar only
    *inertia = t;
  }
  return TM_OK;
}

extern "C" int tm_kmeans_fit(const double *x, int64_t n, int dim, int k, int max_iter, const double *init, uint64_t seed, int nan_empty,
                             int32_t *labels, double *centroids, double *inertia, int *iters) {
  RC(require_gpu());
  if (n < 1 || n > 0x7fffffff || dim < 1 || dim > 1024 || k < 1 || max_iter < 0 || !x || !labels || !centroids)
    return fail(TM_ERR_ARG, "tm_kmeans_fit: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const double *d_x = s.in(x, (size_t)n * dim), *d_init = s.in(init, (size_t)k * dim);
  int32_t *d_labels = s.out(labels, (size_t)n);
  double *d_cent = s.out(centroids, (size_t)k * dim);
  if (s.err == TM_OK) s.err = kmeans_fit_dev(d_x, nullptr, n, dim, k, max_iter, d_init, seed, nan_empty, d_labels, d_cent, nullptr, inertia, iters, s);
  RC(s.finish(true));
  return TM_OK;
}

// ---- int16 points: candidate search on the tensor cores, exact f64 decision (see kmeans.cu)
struct KmI16 {   // per-fit device state
  const int16_t *x; int64_t n; int k;
  uint8_t *x_limbs; uint32_t *x_norm;       // points, split once
  int16_t *rc; uint8_t *c_limbs; uint32_t *c_norm;   // rounded centroids, rebuilt every iteration
  int32_t *cand; uint32_t *cdist; int32_t *amb_list; int32_t *counters;   // counters[0] = changed, [1] = ambiguous
  double *dist;
};
static int km_i16_setup(KmI16 &w, const int16_t *d_x, int64_t n, int k, Stage &s) {
  w.x = d_x; w.n = n; w.k = k;
  const int64_t kpad = (k + 63) / 64 * 64;
  w.x_limbs = (uint8_t *)s.temp((size_t)n * 384); w.x_norm = (uint32_t *)s.temp((size_t)n * 4);
  w.rc = (int16_t *)s.temp((size_t)k * 384); w.c_limbs = (uint8_t *)s.temp((size_t)k * 384); w.c_norm = (uint32_t *)s.temp((size_t)kpad * 4);
  w.cand = (int32_t *)s.temp((size_t)n * KMEANS_KC * 4); w.cdist = (uint32_t *)s.temp((size_t)n * KMEANS_KC * 4);
  w.amb_list = (int32_t *)s.temp((size_t)n * 4); w.counters = (int32_t *)s.temp(8);
  w.dist = (double *)s.temp((size_t)n * 8);
  if (s.err) return s.err;
  if (cudaMemsetAsync(w.c_norm, 0, (size_t)kpad * 4, s.st) != cudaSuccess) return TM_ERR_CUDA;
  return launch_limb_split(d_x, n, w.x_limbs, w.x_norm, s.st);
}
// one assignment pass; *changed / *n_amb are host outputs (synchronises)
static int km_i16_assign(KmI16 &w, const double *d_cent, int32_t *d_labels, int *changed, int *n_amb, Stage &s) {
  int rc = launch_round_centroids(d_cent, w.k, w.rc, s.st);
  if (rc) return rc;
  rc = launch_limb_split(w.rc, w.k, w.c_limbs, w.c_norm, s.st);
  if (rc) return rc;
  const int kc = KMEANS_KC;
  rc = launch_knn_i8(w.x_limbs, w.x_norm, (int)w.n, w.c_limbs, w.c_norm, w.k, kc, w.cand, w.cdist, nullptr, num_sms(), 1, s.st);
  if (rc) return rc;
  if (cudaMemsetAsync(w.counters, 0, 8, s.st) != cudaSuccess) return TM_ERR_CUDA;
  rc = launch_kmeans_rerank_i16(w.x, w.n, w.cand, w.cdist, d_cent, w.k, d_labels, w.dist, w.counters, w.amb_list, w.counters + 1, s.st);
  if (rc) return rc;
  int32_t h[2] = {0, 0};
  if (cudaMemcpyAsync(h, w.counters, 8, cudaMemcpyDeviceToHost, s.st) != cudaSuccess || cudaStreamSynchronize(s.st) != cudaSuccess) return TM_ERR_CUDA;
  int n_bf = 0;
  if (h[1] > 0) {
    // level 2: the uncertified points again on the tensor cores with 64 candidates
    const int na = h[1];
    uint8_t *al = (uint8_t *)s.temp((size_t)na * 384);
    uint32_t *an = (uint32_t *)s.temp((size_t)na * 4), *ad = (uint32_t *)s.temp((size_t)na * 64 * 4);
    int32_t *ac = (int32_t *)s.temp((size_t)na * 64 * 4), *a2 = (int32_t *)s.temp((size_t)na * 4), *a2n = (int32_t *)s.temp(4);
    if (s.err) return s.err;
    if (cudaMemsetAsync(a2n, 0, 4, s.st) != cudaSuccess) return TM_ERR_CUDA;
    rc = launch_amb_gather_limbs(w.x_limbs, w.x_norm, w.amb_list, na, al, an, s.st);
    if (rc) return rc;
    rc = launch_knn_i8(al, an, na, w.c_limbs, w.c_norm, w.k, 64, ac, ad, nullptr, num_sms(), 1, s.st);
    if (rc) return rc;
    rc = launch_kmeans_rerank64(w.x, w.amb_list, na, ac, ad, d_cent, w.k, d_labels, w.dist, w.counters, a2, a2n, s.st);
    if (rc) return rc;
    int32_t h2 = 0;
    if (cudaMemcpyAsync(&h2, a2n, 4, cudaMemcpyDeviceToHost, s.st) != cudaSuccess || cudaStreamSynchronize(s.st) != cudaSuccess) return TM_ERR_CUDA;
    n_bf = h2;
    if (h2 > 0) {   // level 3: exact f64 scan of the residue
      double *xa = (double *)s.temp((size_t)h2 * 192 * 8), *da = (double *)s.temp((size_t)h2 * 8);
      int32_t *la = (int32_t *)s.temp((size_t)h2 * 4);
      if (s.err) return s.err;
      rc = launch_kmeans_assign_amb(w.x, a2, h2, d_cent, w.k, d_labels, w.dist, w.counters, xa, la, da, s.st);
      if (rc) return rc;
    }
    if (cudaMemcpyAsync(h, w.counters, 4, cudaMemcpyDeviceToHost, s.st) != cudaSuccess || cudaStreamSynchronize(s.st) != cudaSuccess) return TM_ERR_CUDA;
  }
  *changed = h[0];
  *n_amb = n_bf;   // points that needed the brute-force scan
  return TM_OK;
}
static double sum_dist(const double *d_dist, int64_t n, cudaStream_t st, int *err) {
  std::vector<double> h((size_t)n);
  if (cudaMemcpyAsync(h.data(), d_dist, (size_t)n * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { *err = TM_ERR_CUDA; return 0.0; }
  double t = 0.0;
  for (double v : h) t += v;   // reporting scalar only
  return t;
}

extern "C" int tm_kmeans_fit_i16(const int16_t *x, int64_t n, int k, int max_iter, const double *init, int nan_empty, int32_t *labels,
                                 double *centroids, double *inertia, int *iters, int64_t *ambiguous) {
  RC(require_gpu());
  if (n < 1 || n > 0x7fffffff || k < 1 || max_iter < 0 || !x || !init || !labels || !centroids)
    return fail(TM_ERR_ARG, "tm_kmeans_fit_i16: bad argument (explicit initial centroids required)");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *d_x = s.in(x, (size_t)n * 192);
  const double *d_init = s.in(init, (size_t)k * 192);
  int32_t *d_labels = s.out(labels, (size_t)n);
  double *d_cent = s.out(centroids, (size_t)k * 192);
  KmI16 w;
  const size_t ws_bytes = kmeans_update_ws_bytes(n, k);
  void *ws = s.temp(ws_bytes);
  int it = 0;
  int64_t amb_total = 0;
  if (s.err == TM_OK) s.err = km_i16_setup(w, d_x, n, k, s);
  if (s.err == TM_OK && cudaMemcpyAsync(d_cent, d_init, (size_t)k * 192 * 8, cudaMemcpyDeviceToDevice, s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
  if (s.err == TM_OK && cudaMemsetAsync(d_labels, 0xFF, (size_t)n * 4, s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
  while (s.err == TM_OK) {
    int changed = 0, n_amb = 0;
    s.err = km_i16_assign(w, d_cent, d_labels, &changed, &n_amb, s);
    amb_total += n_amb;
    if (s.err != TM_OK || changed == 0 || it >= max_iter) break;
    ++it;
    s.err = launch_kmeans_update_i16(d_x, n, d_labels, k, d_cent, nullptr, ws, ws_bytes, nan_empty, 1, s.st);
  }
  if (s.err == TM_OK && inertia) { int e = TM_OK; *inertia = sum_dist(w.dist, n, s.st, &e); s.err = e; }
  if (iters) *iters = it;
  if (ambiguous) *ambiguous = amb_total;
  RC(s.finish(true));
  return TM_OK;
}

// ---- persistent shard state for the multi-GPU Lloyd loop: the points are split into limb rows ONCE, every step only
// re-rounds the centroids, and nothing of a step's result comes back to the host (the one 8-byte read inside the step is
// the count of uncertified points, which decides whether the second-level search is launched at all).
struct tm_kmeans_i16 {
  int64_t n = 0; int k = 0;
  int16_t *x_own = nullptr;     // device copy of host points (null when the caller's rows already live on the device)
  KmI16 w{};
  void *ws = nullptr; size_t ws_bytes = 0;
  double *part = nullptr;       // 1024 block partials of the inertia sum
};

extern "C" int tm_kmeans_i16_destroy(tm_kmeans_i16 *h) {
  if (!h) return TM_OK;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaFree(h->x_own); cudaFree(h->w.x_limbs); cudaFree(h->w.x_norm); cudaFree(h->w.rc); cudaFree(h->w.c_limbs); cudaFree(h->w.c_norm);
  cudaFree(h->w.cand); cudaFree(h->w.cdist); cudaFree(h->w.amb_list); cudaFree(h->w.counters); cudaFree(h->w.dist); cudaFree(h->ws);
  cudaFree(h->part);
  delete h;
  return TM_OK;
}

extern "C" int tm_kmeans_i16_create(const int16_t *x, int64_t n, int k, tm_kmeans_i16 **out) {
  RC(require_gpu());
  if (!x || !out || n < 1 || n > 0x7fffffff || k < 1) return fail(TM_ERR_ARG, "tm_kmeans_i16_create: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaStream_t st = t_stream;
  tm_kmeans_i16 *h = new tm_kmeans_i16();
  h->n = n; h->k = k;
  KmI16 &w = h->w;
  const int64_t kpad = (k + 63) / 64 * 64;
  h->ws_bytes = kmeans_update_ws_bytes(n, k);
  bool ok = true;
  auto alloc = [&](void **p, size_t bytes) { if (ok && cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) ok = false; };
  if (!is_device_ptr(x)) {
    alloc((void **)&h->x_own, (size_t)n * 384);
    if (ok && cudaMemcpyAsync(h->x_own, x, (size_t)n * 384, cudaMemcpyHostToDevice, st) != cudaSuccess) ok = false;
  }
  w.x = h->x_own ? h->x_own : x; w.n = n; w.k = k;
  alloc((void **)&w.x_limbs, (size_t)n * 384); alloc((void **)&w.x_norm, (size_t)n * 4);
  alloc((void **)&w.rc, (size_t)k * 384); alloc((void **)&w.c_limbs, (size_t)k * 384); alloc((void **)&w.c_norm, (size_t)kpad * 4);
  alloc((void **)&w.cand, (size_t)n * KMEANS_KC * 4); alloc((void **)&w.cdist, (size_t)n * KMEANS_KC * 4);
  alloc((void **)&w.amb_list, (size_t)n * 4); alloc((void **)&w.counters, 8); alloc((void **)&w.dist, (size_t)n * 8);
  alloc(&h->ws, h->ws_bytes); alloc((void **)&h->part, 1024 * 8);
  int rc = ok ? TM_OK : TM_ERR_NOMEM;
  if (rc == TM_OK && cudaMemsetAsync(w.c_norm, 0, (size_t)kpad * 4, st) != cudaSuccess) rc = TM_ERR_CUDA;
  if (rc == TM_OK) rc = launch_limb_split(w.x, n, w.x_limbs, w.x_norm, st);
  if (rc == TM_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = TM_ERR_CUDA;
  if (rc != TM_OK) { tm_kmeans_i16_destroy(h); return fail(rc, "tm_kmeans_i16_create"); }
  *out = h;
  return TM_OK;
}

extern "C" int tm_kmeans_i16_step(tm_kmeans_i16 *h, const double *centroids, int32_t *labels, double *partial_sums, int64_t *partial_counts,
                                  int64_t *stats, double *inertia) {
  RC(require_gpu());
  if (!h || !centroids || !labels || !partial_sums || !partial_counts) return fail(TM_ERR_ARG, "tm_kmeans_i16_step: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const double *d_cent = s.in(centroids, (size_t)h->k * 192);
  int32_t *d_labels = s.inout(labels, (size_t)h->n);
  double *d_sums = s.out(partial_sums, (size_t)h->k * 192);
  int64_t *d_counts = s.out(partial_counts, (size_t)h->k);
  int64_t *d_stats = stats ? s.inout(stats, 2) : nullptr;
  double *d_inertia = inertia ? s.out(inertia, 1) : nullptr;
  int ch = 0, n_bf = 0;
  if (s.err == TM_OK) s.err = km_i16_assign(h->w, d_cent, d_labels, &ch, &n_bf, s);
  if (s.err == TM_OK) s.err = launch_kmeans_update_i16(h->w.x, h->n, d_labels, h->k, d_sums, d_counts, h->ws, h->ws_bytes, 0, 0, s.st);
  if (s.err == TM_OK && d_stats) s.err = launch_kmeans_stats_add(d_stats, h->w.counters, n_bf, s.st);
  if (s.err == TM_OK && d_inertia) s.err = launch_sum_f64(h->w.dist, h->n, h->part, d_inertia, s.st);
  RC(s.finish());
  return TM_OK;
}

// multi-GPU building block: one assignment over this rank's shard + per-cluster partial sums/counts
extern "C" int tm_kmeans_partial_step_i16(const int16_t *x, int64_t n, int k, const double *centroids, int32_t *labels,
                                          double *partial_sums, int64_t *partial_counts, int64_t *changed, double *inertia) {
  RC(require_gpu());
  if (n < 1 || n > 0x7fffffff || k < 1 || !x || !centroids || !labels || !partial_sums || !partial_counts)
    return fail(TM_ERR_ARG, "tm_kmeans_partial_step_i16: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *d_x = s.in(x, (size_t)n * 192);
  const double *d_cent = s.in(centroids, (size_t)k * 192);
  int32_t *d_labels = s.inout(labels, (size_t)n);
  double *d_sums = s.out(partial_sums, (size_t)k * 192);
  int64_t *d_counts = s.out(partial_counts, (size_t)k);
  KmI16 w;
  const size_t ws_bytes = kmeans_update_ws_bytes(n, k);
  void *ws = s.temp(ws_bytes);
  int ch = 0, n_amb = 0;
  if (s.err == TM_OK) s.err = km_i16_setup(w, d_x, n, k, s);
  if (s.err == TM_OK) s.err = km_i16_assign(w, d_cent, d_labels, &ch, &n_amb, s);
  if (s.err == TM_OK) s.err = launch_kmeans_update_i16(d_x, n, d_labels, k, d_sums, d_counts, ws, ws_bytes, 0, 0, s.st);
  if (changed) *changed = ch;
  if (s.err == TM_OK && inertia) { int e = TM_OK; *inertia = sum_dist(w.dist, n, s.st, &e); s.err = e; }
  RC(s.finish(true));
  return TM_OK;
}

extern "C" int tm_kmeans_partial_step(const double *x, int64_t n, int dim, int k, const double *centroids, int32_t *labels,
                                      double *partial_sums, int64_t *partial_counts, int64_t *changed, double *inertia) {
  RC(require_gpu());
  if (n < 0 || n > 0x7fffffff || dim < 1 || dim > 1024 || k < 1 || !centroids || !partial_sums || !partial_counts || (n > 0 && (!x || !labels)))
    return fail(TM_ERR_ARG, "tm_kmeans_partial_step: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const double *d_x = s.in(x, (size_t)n * dim), *d_cent = s.in(centroids, (size_t)k * dim);
  int32_t *d_labels = s.inout(labels, (size_t)n);
  double *d_sums = s.out(partial_sums, (size_t)k * dim);
  int64_t *d_counts = s.out(partial_counts, (size_t)k);
  int32_t *d_changed = (int32_t *)s.temp(4);
  double *d_dist = (double *)s.temp((size_t)(n > 0 ? n : 1) * 8);
  const size_t ws_bytes = kmeans_update_ws_bytes(n > 0 ? n : 1, k, dim);
  void *ws = s.temp(ws_bytes);
  int32_t h_changed = 0;
  if (s.err == TM_OK) {
    cudaMemsetAsync(d_changed, 0, 4, s.st);
    if (n > 0) {
      s.err = launch_kmeans_assign_f64(d_x, n, dim, d_cent, k, d_labels, d_dist, d_changed, s.st);
      if (s.err == TM_OK) s.err = launch_kmeans_update_f64(d_x, nullptr, n, dim, d_labels, k, d_sums, d_counts, nullptr, ws, ws_bytes, 0, 0, s.st);
    } else {
      cudaMemsetAsync(d_sums, 0, (size_t)k * dim * 8, s.st);
      cudaMemsetAsync(d_counts, 0, (size_t)k * 8, s.st);
    }
    if (s.err == TM_OK && (changed || inertia)) {
      std::vector<double> h((size_t)n);
      cudaMemcpyAsync(&h_changed, d_changed, 4, cudaMemcpyDeviceToHost, s.st);
      if (inertia && n > 0) cudaMemcpyAsync(h.data(), d_dist, (size_t)n * 8, cudaMemcpyDeviceToHost, s.st);
      if (cudaStreamSynchronize(s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
      if (changed) *changed = h_changed;
      if (inertia) { double t = 0.0; for (double v : h) t += v; *inertia = t; }
    }
  }
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_kmeans_finish_step(const double *sums, const int64_t *counts, int k, int dim, int nan_empty, double *centroids) {
  RC(require_gpu());
  if (k < 1 || dim < 1 || !sums || !counts || !centroids) return fail(TM_ERR_ARG, "tm_kmeans_finish_step: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const double *d_sums = s.in(sums, (size_t)k * dim);
  const int64_t *d_counts = s.in(counts, (size_t)k);
  double *d_cent = s.inout(centroids, (size_t)k * dim);
  if (s.err == TM_OK) s.err = launch_kmeans_finish(d_sums, d_counts, k, dim, nan_empty, d_cent, s.st);
  RC(s.finish());
  return TM_OK;
}

// ------------------------------------------------------------------ palette colour quantisation
extern "C" int tm_palquant_kmeans(const int32_t *rgb, const int32_t *tile_pal, int64_t n_tiles, int n_pal, int pal_size, const double *init,
                                  uint64_t seed, int32_t *palettes, int *iters) {
  RC(require_gpu());
  if (n_tiles < 1 || n_tiles * 64 > 0x7fffffff || n_pal < 1 || pal_size < 1 || pal_size > 256 || !rgb || !tile_pal || !palettes)
    return fail(TM_ERR_ARG, "tm_palquant_kmeans: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n_tiles * 64), *d_tp = s.in(tile_pal, (size_t)n_tiles);
  const double *d_init = s.in(init, (size_t)n_pal * pal_size * 3);
  int32_t *d_out = s.out(palettes, (size_t)n_pal * pal_size);
  int32_t it = 0;
  if (s.err == TM_OK) s.err = run_palette_quantise(d_rgb, d_tp, n_tiles, n_pal, pal_size, d_init, seed, 300, d_out, &it, s.st);
  if (iters) *iters = it;
  RC(s.finish(true));
  return TM_OK;
}

// ------------------------------------------------------------------ dlquant (quantizer.c; extern.pas:195-196)
static int dlquant_batch(int which, const uint8_t *rgb888, const int64_t *img_off, int n_img, int quant_to, int bpc, uint8_t *palettes,
                         int32_t *counts) {
  RC(require_gpu());
  if (n_img < 1 || !rgb888 || !img_off || !palettes || quant_to < 1 || quant_to > 65536 || bpc < 1 || bpc > 5)
    return fail(TM_ERR_ARG, "tm_dlNquant_batch: bad argument (1 <= lookup_bpc <= 5)");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  // image offsets are needed on both sides: bring them to the host if they live on the device
  std::vector<int64_t> h_off((size_t)n_img + 1);
  if (cudaMemcpyAsync(h_off.data(), img_off, h_off.size() * 8, cudaMemcpyDefault, s.st) != cudaSuccess || cudaStreamSynchronize(s.st) != cudaSuccess)
    return fail(TM_ERR_CUDA, "tm_dlNquant_batch: offsets");
  int64_t max_px = 0;
  for (int i = 0; i < n_img; ++i) {
    if (h_off[i + 1] <= h_off[i]) return fail(TM_ERR_ARG, "tm_dlNquant_batch: empty image");
    max_px = std::max(max_px, h_off[i + 1] - h_off[i]);
  }
  const uint8_t *d_rgb = s.in(rgb888, (size_t)h_off[n_img] * 3);
  const int64_t *d_off = s.in(img_off, (size_t)n_img + 1);
  uint8_t *d_pal = s.out(palettes, (size_t)n_img * quant_to * 3);
  int32_t *d_cnt = s.out(counts, (size_t)n_img);
  if (s.err == TM_OK)
    s.err = which == 3 ? run_dl3quant(d_rgb, d_off, n_img, max_px, quant_to, bpc, d_pal, d_cnt, s.st)
                       : run_dl1quant(d_rgb, d_off, n_img, max_px, quant_to, bpc, d_pal, d_cnt, s.st);
  RC(s.finish());
  return TM_OK;
}
extern "C" int tm_dl3quant_batch(const uint8_t *rgb888, const int64_t *img_off, int n_img, int quant_to, int lookup_bpc, uint8_t *palettes,
                                 int32_t *counts) {
  return dlquant_batch(3, rgb888, img_off, n_img, quant_to, lookup_bpc, palettes, counts);
}
extern "C" int tm_dl1quant_batch(const uint8_t *rgb888, const int64_t *img_off, int n_img, int quant_to, int lookup_bpc, uint8_t *palettes,
                                 int32_t *counts) {
  return dlquant_batch(1, rgb888, img_off, n_img, quant_to, lookup_bpc, palettes, counts);
}
// drop-in: int dlNquant(uchar *inbuf, int width, int height, int quant_to, int lookup_bpc, uchar userpal[3][65536])
static int dlquant_dropin(int which, uint8_t *inbuf, int width, int height, int quant_to, int lookup_bpc, uint8_t *userpal) {
  if (!inbuf || !userpal || width < 1 || height < 1 || quant_to < 1 || quant_to > 65536) return 1;
  const int64_t off[2] = {0, (int64_t)width * height};
  std::vector<uint8_t> pal((size_t)quant_to * 3);
  if (dlquant_batch(which, inbuf, off, 1, quant_to, lookup_bpc, pal.data(), nullptr) != TM_OK) return 1;
  memset(userpal, 0, 3 * 65536);   // the DLL's context is calloc'ed: unused entries read 0
  for (int i = 0; i < quant_to; ++i)
    for (int c = 0; c < 3; ++c) userpal[c * 65536 + i] = pal[(size_t)i * 3 + c];
  return 0;
}
extern "C" int dl3quant(uint8_t *inbuf, int width, int height, int quant_to, int lookup_bpc, uint8_t *userpal) {
  return dlquant_dropin(3, inbuf, width, height, quant_to, lookup_bpc, userpal);
}
extern "C" int dl1quant(uint8_t *inbuf, int width, int height, int quant_to, int lookup_bpc, uint8_t *userpal) {
  return dlquant_dropin(1, inbuf, width, height, quant_to, lookup_bpc, userpal);
}

// ------------------------------------------------------------------ matcher
extern "C" int tm_matcher_destroy(tm_matcher *m) {
  if (!m) return TM_OK;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaFreeAsync(m->dict_idx, t_stream); cudaFreeAsync(m->dict_pal, t_stream); cudaFreeAsync(m->palettes, t_stream);
  cudaFreeAsync(m->dict_feat, t_stream); cudaFreeAsync(m->pair_feat, t_stream); cudaFreeAsync(m->pair_norm, t_stream);
  tm_knn_short_destroy(m->knn);
  delete m;
  return TM_OK;
}

extern "C" int tm_matcher_create(const uint8_t *dict_idx, const int32_t *dict_pal, int64_t n_dict, const int32_t *palettes, int pal_size,
                                 int n_pal, int extended, tm_matcher **out) {
  RC(require_gpu());
  if (n_dict < 1 || n_dict > 0x7fffffff || pal_size < 1 || pal_size > 256 || n_pal < 1 || !dict_idx || !dict_pal || !palettes || !out)
    return fail(TM_ERR_ARG, "tm_matcher_create: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaStream_t st = t_stream;
  tm_matcher *m = new tm_matcher();
  m->n_dict = n_dict; m->n_pal = n_pal; m->pal_size = pal_size; m->extended = extended;
  const size_t pair_bytes = extended ? (size_t)n_dict * n_pal * 384 : 0;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  int rc = TM_OK;
  if (pair_bytes + (size_t)n_dict * 1024 > free_b * 9 / 10) rc = TM_ERR_NOMEM;
  if (rc == TM_OK && (cudaMallocAsync((void **)&m->dict_idx, (size_t)n_dict * 64, st) != cudaSuccess ||
                      cudaMallocAsync((void **)&m->dict_pal, (size_t)n_dict * 4, st) != cudaSuccess ||
                      cudaMallocAsync((void **)&m->palettes, (size_t)n_pal * pal_size * 4, st) != cudaSuccess ||
                      cudaMallocAsync((void **)&m->dict_feat, (size_t)n_dict * 384, st) != cudaSuccess ||
                      (extended && (cudaMallocAsync((void **)&m->pair_feat, pair_bytes, st) != cudaSuccess ||
                                    cudaMallocAsync((void **)&m->pair_norm, (size_t)n_dict * n_pal * 4, st) != cudaSuccess))))
    rc = TM_ERR_NOMEM;
  if (rc == TM_OK) {
    if (cudaMemcpyAsync(m->dict_idx, dict_idx, (size_t)n_dict * 64, cudaMemcpyDefault, st) != cudaSuccess ||
        cudaMemcpyAsync(m->dict_pal, dict_pal, (size_t)n_dict * 4, cudaMemcpyDefault, st) != cudaSuccess ||
        cudaMemcpyAsync(m->palettes, palettes, (size_t)n_pal * pal_size * 4, cudaMemcpyDefault, st) != cudaSuccess)
      rc = TM_ERR_CUDA;
  }
  if (rc == TM_OK) rc = launch_features_pal(m->dict_idx, m->dict_pal, m->palettes, pal_size, n_dict, m->dict_feat, st);
  if (rc == TM_OK) rc = knn_short_create_dev(m->dict_feat, n_dict, st, &m->knn);
  if (rc == TM_OK && extended) rc = launch_features_allpairs(m->dict_idx, n_dict, m->palettes, pal_size, n_pal, m->pair_feat, st);
  if (rc == TM_OK && extended) rc = launch_row_norms(m->pair_feat, n_dict * n_pal, m->pair_norm, st);
  if (rc == TM_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = TM_ERR_CUDA;
  if (rc != TM_OK) { tm_matcher_destroy(m); return fail(rc, "tm_matcher_create"); }
  *out = m;
  return TM_OK;
}

static int match_feat_dev(tm_matcher *m, const int16_t *d_feat, int64_t n_q, int k, int32_t *d_tile, int32_t *d_pal, uint32_t *d_err, Stage &s) {
  if (n_q == 0) return TM_OK;
  const int kk = m->extended ? k : 1;
  int32_t *d_idx = (int32_t *)s.temp((size_t)n_q * kk * 4);
  uint32_t *d_dist = (uint32_t *)s.temp((size_t)n_q * kk * 4);
  if (s.err) return s.err;
  int rc = knn_short_batch_dev(m->knn, d_feat, n_q, kk, d_idx, d_dist, 0, s);
  if (rc) return rc;
  if (m->extended)
    return launch_match_rerank(d_feat, n_q, d_idx, kk, m->dict_pal, m->dict_idx, m->n_dict, m->palettes, m->pal_size, m->n_pal, m->pair_feat,
                               m->pair_norm, d_tile, d_pal, d_err, s.st);
  return launch_match_plain(d_idx, d_dist, n_q, m->dict_pal, m->n_dict, d_tile, d_pal, d_err, s.st);
}

extern "C" int tm_match_tiles_feat(tm_matcher *m, const int16_t *feat, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err) {
  RC(require_gpu());
  if (!m || n_q < 0 || n_q > 0x7fffffff || k < 1 || k > 64 || (n_q > 0 && (!feat || !tile_idx || !pal_idx || !err)))
    return fail(TM_ERR_ARG, "tm_match_tiles_feat: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int16_t *d_feat = s.in(feat, (size_t)n_q * 192);
  int32_t *d_tile = s.out(tile_idx, (size_t)n_q), *d_pal = s.out(pal_idx, (size_t)n_q);
  uint32_t *d_err = s.out(err, (size_t)n_q);
  if (s.err == TM_OK) s.err = match_feat_dev(m, d_feat, n_q, k, d_tile, d_pal, d_err, s);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_match_tiles_rgb(tm_matcher *m, const int32_t *rgb, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err) {
  RC(require_gpu());
  if (!m || n_q < 0 || n_q > 0x7fffffff || k < 1 || k > 64 || (n_q > 0 && (!rgb || !tile_idx || !pal_idx || !err)))
    return fail(TM_ERR_ARG, "tm_match_tiles_rgb: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  int32_t *d_tile = s.out(tile_idx, (size_t)n_q), *d_pal = s.out(pal_idx, (size_t)n_q);
  uint32_t *d_err = s.out(err, (size_t)n_q);
  int16_t *d_feat = (int16_t *)s.temp((size_t)n_q * 384);
  // A large batch of HOST tiles is uploaded in pieces on a second stream while the pieces already on the device are
  // matched: only the first piece (two k-NN waves, 10 MB of the 110 MB of a 720p sequence) stays exposed.  Pieces are
  // whole waves of k-NN query blocks (SMs x 128 rows) so the split costs the search no tail.
  const int64_t wave = (int64_t)num_sms() * knn_rows_per_cta();
  if (!is_device_ptr(rgb) && n_q >= 8 * wave) {
    static cudaStream_t cs_dev[64] = {};        // copy stream and events of each device (one process may drive several)
    static cudaEvent_t ev_dev[64][5] = {};
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(TM_ERR_ARG, "tm_match_tiles_rgb: device ordinal out of range");
    if (!cs_dev[dev]) {
      CU(cudaStreamCreateWithFlags(&cs_dev[dev], cudaStreamNonBlocking));
      for (auto &e : ev_dev[dev]) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t cs = cs_dev[dev];
    cudaEvent_t *ev = ev_dev[dev];
    int32_t *d_rgb = (int32_t *)s.temp((size_t)n_q * 256);
    s.any_host = true;
    // TM_MATCH_PIECES = 2..4: the first piece + 1..3 equal ones.  Default 3: every piece boundary drains the persistent k-NN kernel
    // once (measured end to end on 432 000 tiles: 4 pieces 1.11-1.13e12, 3 pieces 1.135e12, 2 pieces 1.136e12 evals/s), and with
    // 3 the second upload (29 MB) still hides behind the first piece on a slower host link
    static int n_pieces = -1;
    if (n_pieces < 0) { const char *e = getenv("TM_MATCH_PIECES"); n_pieces = e ? atoi(e) : 3; if (n_pieces < 2 || n_pieces > 4) n_pieces = 3; }
    const int rest = n_pieces - 1;
    const int64_t first = 2 * wave, piece = ((n_q - first + rest - 1) / rest + wave - 1) / wave * wave;
    int64_t cut[5] = {0, first, n_q, n_q, n_q};
    for (int i = 2; i < 4; ++i) if (i <= rest) cut[i] = first + (i - 1) * piece < n_q ? first + (i - 1) * piece : n_q;
    if (s.err == TM_OK) {
      // Copy i + 1 is queued AFTER the kernels of piece i are launched: with pageable input the host blocks inside
      // cudaMemcpyAsync while it stages the bytes, and this order lets the resident piece be matched meanwhile.  No early
      // return below: every exit first drains the copy stream, so neither the scratch (freed on s.st) nor the caller's host
      // buffer is still in use by a copy when the call returns.
      auto ok = [&](cudaError_t e) { if (e != cudaSuccess && s.err == TM_OK) s.err = TM_ERR_CUDA; return e == cudaSuccess; };
      auto queue_copy = [&](int i) {
        const int64_t off = cut[i], n = cut[i + 1] - off;
        if (n > 0) ok(cudaMemcpyAsync(d_rgb + off * 64, rgb + off * 64, (size_t)n * 256, cudaMemcpyHostToDevice, cs));
        ok(cudaEventRecord(ev[i], cs));
      };
      ok(cudaEventRecord(ev[4], s.st));   // the scratch (stream-ordered allocation on s.st) exists
      ok(cudaStreamWaitEvent(cs, ev[4], 0));
      if (s.err == TM_OK) queue_copy(0);
      for (int i = 0; i < 4 && s.err == TM_OK; ++i) {
        const int64_t off = cut[i], n = cut[i + 1] - off;
        ok(cudaStreamWaitEvent(s.st, ev[i], 0));
        if (n > 0 && s.err == TM_OK) s.err = launch_features_rgb(d_rgb + off * 64, n, d_feat + off * 192, s.st);
        if (n > 0 && s.err == TM_OK) s.err = match_feat_dev(m, d_feat + off * 192, n, k, d_tile + off, d_pal + off, d_err + off, s);
        if (i + 1 < 4 && s.err == TM_OK) queue_copy(i + 1);
      }
      if (s.err != TM_OK) cudaStreamSynchronize(cs);
    }
    RC(s.finish());
    return TM_OK;
  }
  const int32_t *d_rgb = s.in(rgb, (size_t)n_q * 64);
  if (s.err == TM_OK) s.err = launch_features_rgb(d_rgb, n_q, d_feat, s.st);
  if (s.err == TM_OK) s.err = match_feat_dev(m, d_feat, n_q, k, d_tile, d_pal, d_err, s);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_features_from_rgb_mirrored(const int32_t *rgb, const uint8_t *flags, int64_t n, int16_t *out) {
  RC(require_gpu());
  if (n < 0 || (n > 0 && (!rgb || !flags || !out))) return fail(TM_ERR_ARG, "tm_features_from_rgb_mirrored: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n * 64);
  const uint8_t *d_fl = s.in(flags, (size_t)n);
  int16_t *d_out = s.out(out, (size_t)n * 192);
  if (s.err == TM_OK) s.err = launch_features_rgb_mirrored(d_rgb, d_fl, n, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

// DoXY's k-NN branch with the 4 mirror variants of every source tile (a superset of the reference: variant 0 is its result)
extern "C" int tm_match_tiles_rgb_mirrors(tm_matcher *m, const int32_t *rgb, int64_t n_q, int k, int32_t *tile_idx, int32_t *pal_idx, uint32_t *err,
                                          uint8_t *variant) {
  RC(require_gpu());
  if (!m || n_q < 0 || n_q * 4 > 0x7fffffff || k < 1 || k > 64 || (n_q > 0 && (!rgb || !tile_idx || !pal_idx || !err || !variant)))
    return fail(TM_ERR_ARG, "tm_match_tiles_rgb_mirrors: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n_q * 64);
  int32_t *d_tile = s.out(tile_idx, (size_t)n_q), *d_pal = s.out(pal_idx, (size_t)n_q);
  uint32_t *d_err = s.out(err, (size_t)n_q);
  uint8_t *d_var = s.out(variant, (size_t)n_q);
  if (n_q > 0) {
    int16_t *d_feat = (int16_t *)s.temp((size_t)n_q * 4 * 384);
    uint8_t *d_fl = (uint8_t *)s.temp((size_t)n_q);
    int32_t *t4 = (int32_t *)s.temp((size_t)n_q * 16), *p4 = (int32_t *)s.temp((size_t)n_q * 16);
    uint32_t *e4 = (uint32_t *)s.temp((size_t)n_q * 16);
    for (int v = 0; v < 4 && s.err == TM_OK; ++v) {   // features of the tile as mirrored by v (bit 0 H, bit 1 V), exact reference arithmetic
      if (cudaMemsetAsync(d_fl, v, (size_t)n_q, s.st) != cudaSuccess) { s.err = TM_ERR_CUDA; break; }
      s.err = launch_features_rgb_mirrored(d_rgb, d_fl, n_q, d_feat + (size_t)v * n_q * 192, s.st);
    }
    if (s.err == TM_OK) s.err = match_feat_dev(m, d_feat, n_q * 4, k, t4, p4, e4, s);   // one batched search over all variants
    if (s.err == TM_OK) s.err = launch_mirror_combine(t4, p4, e4, n_q, d_tile, d_pal, d_err, d_var, s.st);
  }
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_matcher_dict_features(tm_matcher *m, int16_t *out) {
  RC(require_gpu());
  if (!m || !out) return fail(TM_ERR_ARG, "tm_matcher_dict_features: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  CU(cudaMemcpyAsync(out, m->dict_feat, (size_t)m->n_dict * 384, cudaMemcpyDefault, t_stream));
  CU(cudaStreamSynchronize(t_stream));
  return TM_OK;
}

// ------------------------------------------------------------------ motion search + Reconstruct
// The window scan runs on the tensor cores (motion_tc.cu); TM_MOTION_SCALAR=1 selects the CUDA-core kernel (motion.cu)
// for A/B comparisons.  Both are bit-exact against the oracle.
static int motion_search_dev(const int16_t *d_cur, int tw, int th, const int16_t *d_dcts, int radius, int32_t *d_px, int32_t *d_py,
                             uint32_t *d_err, void *ws, size_t wsb, cudaStream_t st) {
  static int scalar = -1;
  if (scalar < 0) scalar = getenv("TM_MOTION_SCALAR") && atoi(getenv("TM_MOTION_SCALAR")) ? 1 : 0;
  if (scalar) return launch_motion_search(d_cur, tw, th, d_dcts, radius, d_px, d_py, d_err, st);
  int ctas = num_sms();
  if (const char *e = getenv("TM_MOTION_CTAS")) { const int v = atoi(e); if (v > 0 && v < ctas) ctas = v; }   // tests: several tile blocks per CTA
  return launch_motion_search_tc(d_cur, tw, th, d_dcts, radius, d_px, d_py, d_err, ws, wsb, ctas, st);
}

// DoDCTs of `frame` + the window scan of every tile of the current frame.  In the fast feature mode with the tensor-core search the
// sliding-window kernel writes the search's candidate operands (limb rows + norms) itself: no int16 intermediate, no split pass.
// d_dcts: scratch for the int16 features of the other paths ([(w - 7) * (h - 7)][192]).
static int sliding_motion_search_dev(const int32_t *d_frame, int w, int h, int16_t *d_dcts, const int16_t *d_cur, int tw, int th, int radius,
                                     int32_t *d_px, int32_t *d_py, uint32_t *d_err, void *ws, size_t wsb, cudaStream_t st) {
  static int scalar = -1;
  if (scalar < 0) scalar = getenv("TM_MOTION_SCALAR") && atoi(getenv("TM_MOTION_SCALAR")) ? 1 : 0;
  if (!scalar && get_feature_mode() == 1) {
    uint8_t *c_limbs; uint32_t *c_norm; int pwp;
    motion_tc_cand_layout(ws, tw, th, &c_limbs, &c_norm, &pwp);
    int rc = launch_features_sliding_limbs(d_frame, w, h, c_limbs, c_norm, pwp, st);
    if (rc != TM_OK) return rc;
    return motion_search_dev(d_cur, tw, th, nullptr, radius, d_px, d_py, d_err, ws, wsb, st);
  }
  int rc = launch_features_sliding(d_frame, w, h, d_dcts, st);
  if (rc != TM_OK) return rc;
  return motion_search_dev(d_cur, tw, th, d_dcts, radius, d_px, d_py, d_err, ws, wsb, st);
}

extern "C" int tm_sliding_features(const int32_t *frame, int w, int h, int16_t *out) {
  RC(require_gpu());
  if (!frame || !out || w < 8 || h < 8) return fail(TM_ERR_ARG, "tm_sliding_features: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_frame = s.in(frame, (size_t)w * h);
  int16_t *d_out = s.out(out, (size_t)(w - 7) * (h - 7) * 192);
  if (s.err == TM_OK) s.err = launch_features_sliding(d_frame, w, h, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_motion_search(const int16_t *cur_feat, int tw, int th, const int16_t *dcts, int radius, int32_t *pred_x, int32_t *pred_y,
                                uint32_t *err) {
  RC(require_gpu());
  if (!cur_feat || !dcts || !pred_x || !pred_y || !err || tw < 1 || th < 1 || radius < 1 || radius > 128)
    return fail(TM_ERR_ARG, "tm_motion_search: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const size_t nt = (size_t)tw * th;
  const int16_t *d_cur = s.in(cur_feat, nt * 192);
  const int16_t *d_dcts = s.in(dcts, (size_t)(tw * 8 - 7) * (th * 8 - 7) * 192);
  int32_t *d_px = s.out(pred_x, nt), *d_py = s.out(pred_y, nt);
  uint32_t *d_err = s.out(err, nt);
  const size_t wsb = motion_tc_ws_bytes(tw, th);
  void *ws = s.temp(wsb);
  if (s.err == TM_OK) s.err = motion_search_dev(d_cur, tw, th, d_dcts, radius, d_px, d_py, d_err, ws, wsb, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_predict_motion_frame(const int32_t *prev_frame, const int32_t *canon_tiles, const uint8_t *flags, int tw, int th, int radius,
                                       int32_t *pred_x, int32_t *pred_y, uint32_t *err) {
  RC(require_gpu());
  if (!prev_frame || !canon_tiles || !flags || !pred_x || !pred_y || !err || tw < 1 || th < 1 || radius < 1 || radius > 128)
    return fail(TM_ERR_ARG, "tm_predict_motion_frame: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const size_t nt = (size_t)tw * th;
  const int w = tw * 8, h = th * 8;
  const int32_t *d_prev = s.in(prev_frame, (size_t)w * h);
  const int32_t *d_tiles = s.in(canon_tiles, nt * 64);
  const uint8_t *d_flags = s.in(flags, nt);
  int32_t *d_px = s.out(pred_x, nt), *d_py = s.out(pred_y, nt);
  uint32_t *d_err = s.out(err, nt);
  int16_t *d_cur = (int16_t *)s.temp(nt * 384);
  int16_t *d_dcts = (int16_t *)s.temp((size_t)(w - 7) * (h - 7) * 384);
  if (s.err == TM_OK) s.err = launch_features_rgb_mirrored(d_tiles, d_flags, (int64_t)nt, d_cur, s.st);
  const size_t wsb = motion_tc_ws_bytes(tw, th);
  void *ws = s.temp(wsb);
  if (s.err == TM_OK) s.err = sliding_motion_search_dev(d_prev, w, h, d_dcts, d_cur, tw, th, radius, d_px, d_py, d_err, ws, wsb, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_reconstruct_sequence(tm_matcher *m, const int32_t *canon_tiles, const uint8_t *flags, int n_frames, int tw, int th, int radius,
                                       int k, int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x, int32_t *pred_y, uint8_t *is_pred,
                                       uint32_t *err, float *psnr, int32_t *recon) {
  RC(require_gpu());
  if (!m || !canon_tiles || !flags || n_frames < 1 || tw < 1 || th < 1 || radius < 0 || radius > 128 || k < 1 || k > 64 || !tile_idx ||
      !pal_idx || !pred_x || !pred_y || !is_pred || !err)
    return fail(TM_ERR_ARG, "tm_reconstruct_sequence: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const size_t nt = (size_t)tw * th, n_all = nt * (size_t)n_frames;
  const int w = tw * 8, h = th * 8;
  const size_t fpx = (size_t)w * h;
  if (n_all > 0x7fffffff) return fail(TM_ERR_ARG, "tm_reconstruct_sequence: too many tiles in one call");
  const int32_t *d_tiles = s.in(canon_tiles, n_all * 64);
  const uint8_t *d_flags = s.in(flags, n_all);
  int32_t *d_tile = s.out(tile_idx, n_all), *d_pal = s.out(pal_idx, n_all), *d_px = s.out(pred_x, n_all), *d_py = s.out(pred_y, n_all);
  uint8_t *d_isp = s.out(is_pred, n_all);
  uint32_t *d_err = s.out(err, n_all);
  float *d_psnr = psnr ? s.out(psnr, n_all) : nullptr;
  int32_t *d_recon = recon ? s.out(recon, fpx * n_frames) : nullptr;
  int32_t *pp[2] = {nullptr, nullptr};
  if (!d_recon) { pp[0] = (int32_t *)s.temp(fpx * 4); pp[1] = (int32_t *)s.temp(fpx * 4); }
  const bool motion = radius - 1 >= 0 && n_frames > 1;
  int16_t *d_ft = (int16_t *)s.temp(n_all * 384);
  int16_t *d_cur = motion ? (int16_t *)s.temp(n_all * 384) : nullptr;
  int16_t *d_dcts = motion ? (int16_t *)s.temp((size_t)(w - 7) * (h - 7) * 384) : nullptr;
  int32_t *k_tile = (int32_t *)s.temp(n_all * 4), *k_pal = (int32_t *)s.temp(n_all * 4);
  uint32_t *k_err = (uint32_t *)s.temp(n_all * 4);
  int32_t *mx = (int32_t *)s.temp(nt * 4), *my = (int32_t *)s.temp(nt * 4);
  uint32_t *me = (uint32_t *)s.temp(nt * 4);
  const size_t wsb = motion ? motion_tc_ws_bytes(tw, th) : 0;
  void *ws = motion ? s.temp(wsb) : nullptr;
  // the k-NN + re-rank candidates do not depend on the reconstructed frames: one batched pass over the whole sequence
  if (s.err == TM_OK) s.err = launch_features_rgb(d_tiles, (int64_t)n_all, d_ft, s.st);
  if (s.err == TM_OK) s.err = match_feat_dev(m, d_ft, (int64_t)n_all, k, k_tile, k_pal, k_err, s);
  if (s.err == TM_OK && motion) s.err = launch_features_rgb_mirrored(d_tiles, d_flags, (int64_t)n_all, d_cur, s.st);
  for (int f = 0; f < n_frames && s.err == TM_OK; ++f) {
    int32_t *front = d_recon ? d_recon + fpx * f : pp[(f + 1) & 1];
    const int32_t *back = d_recon ? (f > 0 ? d_recon + fpx * (f - 1) : nullptr) : pp[f & 1];
    const bool mo = motion && f > 0;
    if (mo) {
      s.err = sliding_motion_search_dev(back, w, h, d_dcts, d_cur + nt * 192 * f, tw, th, radius, mx, my, me, ws, wsb, s.st);
    }
    const size_t o = nt * f;
    if (s.err == TM_OK)
      s.err = launch_reconstruct_decide(d_flags + o, tw, th, mo ? mx : nullptr, mo ? my : nullptr, mo ? me : nullptr, k_tile + o, k_pal + o,
                                        k_err + o, m->dict_idx, m->palettes, m->pal_size, back, front, d_tile + o, d_pal + o, d_px + o,
                                        d_py + o, d_isp + o, d_err + o, d_psnr ? d_psnr + o : nullptr, s.st);
  }
  RC(s.finish());
  return TM_OK;
}

// One frame of TFrame.Reconstruct for hosts that keep their own frame loop: back = previous reconstructed frame buffer
// (NULL on the first frame of a keyframe sequence: no motion search), front = this frame's reconstruction (written).
extern "C" int tm_reconstruct_frame(tm_matcher *m, const int32_t *canon_tiles, const uint8_t *flags, int tw, int th, int radius, int k,
                                    const int32_t *back, int32_t *front, int32_t *tile_idx, int32_t *pal_idx, int32_t *pred_x,
                                    int32_t *pred_y, uint8_t *is_pred, uint32_t *err, float *psnr) {
  RC(require_gpu());
  if (!m || !canon_tiles || !flags || tw < 1 || th < 1 || radius < 0 || radius > 128 || k < 1 || k > 64 || !front || !tile_idx || !pal_idx ||
      !pred_x || !pred_y || !is_pred || !err)
    return fail(TM_ERR_ARG, "tm_reconstruct_frame: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const size_t nt = (size_t)tw * th;
  const int w = tw * 8, h = th * 8;
  const size_t fpx = (size_t)w * h;
  const bool motion = back != nullptr && radius - 1 >= 0;
  const int32_t *d_tiles = s.in(canon_tiles, nt * 64);
  const uint8_t *d_flags = s.in(flags, nt);
  const int32_t *d_back = motion ? s.in(back, fpx) : nullptr;
  int32_t *d_front = s.out(front, fpx);
  int32_t *d_tile = s.out(tile_idx, nt), *d_pal = s.out(pal_idx, nt), *d_px = s.out(pred_x, nt), *d_py = s.out(pred_y, nt);
  uint8_t *d_isp = s.out(is_pred, nt);
  uint32_t *d_err = s.out(err, nt);
  float *d_psnr = psnr ? s.out(psnr, nt) : nullptr;
  int16_t *d_ft = (int16_t *)s.temp(nt * 384);
  int32_t *k_tile = (int32_t *)s.temp(nt * 4), *k_pal = (int32_t *)s.temp(nt * 4);
  uint32_t *k_err = (uint32_t *)s.temp(nt * 4);
  int32_t *mx = nullptr, *my = nullptr; uint32_t *me = nullptr;
  if (s.err == TM_OK) s.err = launch_features_rgb(d_tiles, (int64_t)nt, d_ft, s.st);
  if (s.err == TM_OK) s.err = match_feat_dev(m, d_ft, (int64_t)nt, k, k_tile, k_pal, k_err, s);
  if (motion && s.err == TM_OK) {
    int16_t *d_cur = (int16_t *)s.temp(nt * 384);
    int16_t *d_dcts = (int16_t *)s.temp((size_t)(w - 7) * (h - 7) * 384);
    const size_t wsb = motion_tc_ws_bytes(tw, th);
    void *ws = s.temp(wsb);
    mx = (int32_t *)s.temp(nt * 4); my = (int32_t *)s.temp(nt * 4); me = (uint32_t *)s.temp(nt * 4);
    if (s.err == TM_OK) s.err = launch_features_rgb_mirrored(d_tiles, d_flags, (int64_t)nt, d_cur, s.st);
    if (s.err == TM_OK) s.err = sliding_motion_search_dev(d_back, w, h, d_dcts, d_cur, tw, th, radius, mx, my, me, ws, wsb, s.st);
  }
  if (s.err == TM_OK)
    s.err = launch_reconstruct_decide(d_flags, tw, th, mx, my, me, k_tile, k_pal, k_err, m->dict_idx, m->palettes, m->pal_size, d_back, d_front,
                                      d_tile, d_pal, d_px, d_py, d_isp, d_err, d_psnr, s.st);
  RC(s.finish());
  return TM_OK;
}

/* mean squared error over the three colour channels of two packed-RGB buffers */
extern "C" int tm_mse_rgb(const int32_t *a, const int32_t *b, int64_t n, double *mse) {
  RC(require_gpu());
  if (!a || !b || n < 1 || !mse) return fail(TM_ERR_ARG, "tm_mse_rgb: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *da = s.in(a, (size_t)n), *db = s.in(b, (size_t)n);
  unsigned long long *acc = (unsigned long long *)s.temp(8);
  unsigned long long host = 0;
  if (s.err == TM_OK && cudaMemsetAsync(acc, 0, 8, s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
  if (s.err == TM_OK) s.err = launch_sq_err_rgb(da, db, n, acc, s.st);
  if (s.err == TM_OK && cudaMemcpyAsync(&host, acc, 8, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
  RC(s.finish(true));
  *mse = (double)host / (3.0 * (double)n);
  return TM_OK;
}

// ------------------------------------------------------------------ Reduce: exact duplicate classes of RGB tiles
extern "C" int tm_tile_classes(const int32_t *rgb, int64_t n, int32_t *class_id, int32_t *n_classes) {
  RC(require_gpu());
  if (!rgb || !class_id || !n_classes || n < 1 || n > 0x7fffffff) return fail(TM_ERR_ARG, "tm_tile_classes: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_rgb = s.in(rgb, (size_t)n * 64);
  int32_t *d_cls = s.out(class_id, (size_t)n);
  const size_t wsb = tile_classes_ws_bytes(n);
  void *ws = s.temp(wsb);
  int32_t *d_cnt = (int32_t *)s.temp(4);
  int32_t host_cnt = 0;
  if (s.err == TM_OK) s.err = run_tile_classes(d_rgb, n, d_cls, d_cnt, ws, wsb, s.st);
  if (s.err == TM_OK && cudaMemcpyAsync(&host_cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s.st) != cudaSuccess) s.err = TM_ERR_CUDA;
  RC(s.finish(true));
  *n_classes = host_cnt;
  return TM_OK;
}

// ---- the tile-count search's bookkeeping over the duplicate classes (STCGREval / TransferTiles, :4014-4103)
extern "C" int tm_reduce_class_min(const int32_t *class_id, const double *eff_psnr, int64_t n, int64_t n_classes, double *sorted_min) {
  RC(require_gpu());
  if (!class_id || !eff_psnr || !sorted_min || n < 1 || n > 0x7fffffff || n_classes < 1 || n_classes > n)
    return fail(TM_ERR_ARG, "tm_reduce_class_min: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_cls = s.in(class_id, (size_t)n);
  const double *d_eff = s.in(eff_psnr, (size_t)n);
  double *d_out = s.out(sorted_min, (size_t)n_classes);
  const size_t wsb = class_min_ws_bytes(n_classes);
  void *ws = s.temp(wsb);
  if (s.err == TM_OK) s.err = run_class_min_sorted(d_cls, d_eff, n, n_classes, d_out, ws, wsb, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_reduce_apply(const int32_t *class_id, const double *eff_psnr, int64_t n, int64_t n_classes, double x, int32_t *use_count,
                               int32_t *first_member, uint8_t *unpredicted) {
  RC(require_gpu());
  if (!class_id || !eff_psnr || !use_count || !first_member || !unpredicted || n < 1 || n > 0x7fffffff || n_classes < 1)
    return fail(TM_ERR_ARG, "tm_reduce_apply: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_cls = s.in(class_id, (size_t)n);
  const double *d_eff = s.in(eff_psnr, (size_t)n);
  int32_t *d_use = s.out(use_count, (size_t)n_classes), *d_rep = s.out(first_member, (size_t)n_classes);
  uint8_t *d_un = s.out(unpredicted, (size_t)n);
  if (s.err == TM_OK) s.err = run_reduce_apply(d_cls, d_eff, n, n_classes, x, d_use, d_rep, d_un, s.st);
  RC(s.finish());
  return TM_OK;
}

extern "C" int tm_reduce_remap(const int32_t *class_id, const uint8_t *unpredicted, const int32_t *new_of_class, int64_t n, int64_t n_classes,
                               int32_t *tile_idx) {
  RC(require_gpu());
  if (!class_id || !unpredicted || !new_of_class || !tile_idx || n < 1 || n_classes < 1) return fail(TM_ERR_ARG, "tm_reduce_remap: bad argument");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  const int32_t *d_cls = s.in(class_id, (size_t)n);
  const uint8_t *d_un = s.in(unpredicted, (size_t)n);
  const int32_t *d_new = s.in(new_of_class, (size_t)n_classes);
  int32_t *d_out = s.out(tile_idx, (size_t)n);
  if (s.err == TM_OK) s.err = run_reduce_remap(d_cls, d_un, d_new, n, d_out, s.st);
  RC(s.finish());
  return TM_OK;
}

// ================================================================== drop-in exports (extern.pas:178-223)
extern "C" tm_knn_short *ann_kdtree_short_create(int16_t **rows, int n, int dim, int bucket, int split) {
  (void)bucket; (void)split;
  if (!rows || n < 1 || dim != 192) { fail(TM_ERR_ARG, "ann_kdtree_short_create: dim must be 192"); return nullptr; }
  std::vector<int16_t> flat((size_t)n * 192);
  for (int i = 0; i < n; ++i) memcpy(&flat[(size_t)i * 192], rows[i], 192 * sizeof(int16_t));   // marshal the row-pointer array
  tm_knn_short *h = nullptr;
  if (tm_knn_short_create(flat.data(), n, &h) != TM_OK) return nullptr;
  return h;
}
extern "C" void ann_kdtree_short_destroy(tm_knn_short *h) { tm_knn_short_destroy(h); }
// one batched search for every request queued on the handle (requests of one batch may ask for different k: grouped)
static void knn_short_exec(tm_knn_short *h, std::vector<KnnReq *> &batch) {
  std::vector<bool> taken(batch.size(), false);
  for (size_t a = 0; a < batch.size(); ++a) {
    if (taken[a]) continue;
    const int k = batch[a]->k;
    std::vector<size_t> grp;
    for (size_t b = a; b < batch.size(); ++b)
      if (!taken[b] && batch[b]->k == k) { grp.push_back(b); taken[b] = true; }
    const size_t m = grp.size();
    std::vector<int16_t> qbuf(m * 192);
    std::vector<int32_t> ibuf(m * k);
    std::vector<uint32_t> dbuf(m * k);
    for (size_t i = 0; i < m; ++i) memcpy(&qbuf[i * 192], batch[grp[i]]->q, 192 * sizeof(int16_t));
    const bool ok = k >= 1 && k <= 64 && tm_knn_short_batch(h, qbuf.data(), (int64_t)m, k, ibuf.data(), dbuf.data(), 1) == TM_OK;
    for (size_t i = 0; i < m; ++i) {
      KnnReq *r = batch[grp[i]];
      for (int j = 0; j < k; ++j) {   // out-of-range = "no result" for the host (:1549, :1566)
        r->idx[j] = ok ? ibuf[i * k + j] : -1;
        if (r->dist) ((uint32_t *)r->dist)[j] = ok ? dbuf[i * k + j] : 0xFFFFFFFFu;
      }
    }
  }
}
extern "C" int ann_kdtree_short_search(tm_knn_short *h, const int16_t *q, uint32_t eps, uint32_t *err) {
  (void)eps;
  int32_t idx = -1; uint32_t d = 0xFFFFFFFFu;
  if (h && q) {
    KnnReq r{q, 1, &idx, &d, false};
    h->rv.submit(r, [h](std::vector<KnnReq *> &b) { knn_short_exec(h, b); });
  }
  if (err) *err = d;
  return idx;
}
extern "C" void ann_kdtree_short_search_multi(tm_knn_short *h, int *idxs, uint32_t *errs, int k, const int16_t *q, uint32_t eps) {
  (void)eps;
  if (k < 1 || !idxs || !errs) return;
  if (!h || !q || k > 64) {
    for (int i = 0; i < k; ++i) { idxs[i] = -1; errs[i] = 0xFFFFFFFFu; }
    return;
  }
  KnnReq r{q, k, idxs, errs, false};
  h->rv.submit(r, [h](std::vector<KnnReq *> &b) { knn_short_exec(h, b); });
}
/* queries answered / batched launches so far on this handle (tests, tuning) */
extern "C" int tm_rendezvous_stats(tm_knn_short *h, int64_t *queries, int64_t *batches) {
  if (!h) return fail(TM_ERR_ARG, "tm_rendezvous_stats: bad argument");
  std::lock_guard<std::mutex> lk(h->rv.mu);
  if (queries) *queries = h->rv.queries;
  if (batches) *batches = h->rv.batches;
  return TM_OK;
}

extern "C" tm_knn_double *ann_kdtree_create(double **rows, int n, int dim, int bucket, int split) {
  (void)bucket; (void)split;
  if (require_gpu() != TM_OK) return nullptr;
  if (!rows || n < 1 || dim < 1) { fail(TM_ERR_ARG, "ann_kdtree_create: bad argument"); return nullptr; }
  std::vector<double> flat((size_t)n * dim);
  for (int i = 0; i < n; ++i) memcpy(&flat[(size_t)i * dim], rows[i], (size_t)dim * sizeof(double));
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  tm_knn_double *h = new tm_knn_double();
  h->n = n; h->dim = dim;
  if (cudaMalloc(&h->pts, flat.size() * 8) != cudaSuccess || cudaMemcpy(h->pts, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(h->pts); delete h; fail(TM_ERR_NOMEM, "ann_kdtree_create"); return nullptr;
  }
  return h;
}
extern "C" void ann_kdtree_destroy(tm_knn_double *h) {
  if (!h) return;
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  cudaFree(h->pts); delete h;
}
extern "C" int ann_kdtree_search(tm_knn_double *h, const double *q, double eps, double *err) {
  (void)eps;
  int32_t idx = -1; double d = INFINITY;
  if (h && q) {
    KnnReq r{q, 1, &idx, &d, false};
    h->rv.submit(r, [h](std::vector<KnnReq *> &batch) {   // DoANN (:4128) calls this once per dictionary tile from the pool threads
      const size_t m = batch.size(), dim = (size_t)h->dim;
      std::vector<double> qbuf(m * dim), dbuf(m);
      std::vector<int32_t> ibuf(m);
      for (size_t i = 0; i < m; ++i) memcpy(&qbuf[i * dim], batch[i]->q, dim * sizeof(double));
      const bool ok = tm_knn_double_batch(h->pts, h->n, h->dim, qbuf.data(), (int64_t)m, ibuf.data(), dbuf.data()) == TM_OK;
      for (size_t i = 0; i < m; ++i) { batch[i]->idx[0] = ok ? ibuf[i] : -1; *(double *)batch[i]->dist = ok ? dbuf[i] : INFINITY; }
    });
  }
  if (err) *err = d;
  return idx;
}

extern "C" tm_yakmo *yakmo_create(uint32_t k, uint32_t restarts, int max_iter, int init_type, int init_seed, int normalize, int verbose) {
  (void)restarts; (void)init_type; (void)normalize; (void)verbose;
  tm_yakmo *y = new tm_yakmo();
  y->k = k; y->max_iter = max_iter; y->seed = (uint64_t)(uint32_t)init_seed;
  return y;
}
extern "C" void yakmo_destroy(tm_yakmo *y) { delete y; }
extern "C" void yakmo_set_num_threads(int n) { (void)n; }
extern "C" void yakmo_load_train_data(tm_yakmo *y, uint32_t rows, uint32_t cols, double **data) {
  if (!y || !data) return;
  y->rows = rows; y->cols = cols;
  y->data.resize((size_t)rows * cols);
  for (uint32_t i = 0; i < rows; ++i) memcpy(&y->data[(size_t)i * cols], data[i], (size_t)cols * sizeof(double));   // the host frees its rows right after (:4495)
}
extern "C" void yakmo_train_on_data(tm_yakmo *y, int *point_to_cluster) {
  if (!y || !point_to_cluster || y->rows == 0) return;
  y->cent.assign((size_t)y->k * y->cols, NAN);
  const int k = (int)(y->k < y->rows ? y->k : y->rows);
  std::vector<double> cent((size_t)k * y->cols);
  if (tm_kmeans_fit(y->data.data(), y->rows, (int)y->cols, k, y->max_iter, nullptr, y->seed, 1, point_to_cluster, cent.data(), nullptr, nullptr) != TM_OK) {
    for (uint32_t i = 0; i < y->rows; ++i) point_to_cluster[i] = 0;
    return;
  }
  memcpy(y->cent.data(), cent.data(), cent.size() * sizeof(double));
}
extern "C" void yakmo_get_centroids(tm_yakmo *y, double **centroids) {
  if (!y || !centroids) return;
  for (uint32_t c = 0; c < y->k; ++c)
    if (centroids[c]) memcpy(centroids[c], &y->cent[(size_t)c * y->cols], (size_t)y->cols * sizeof(double));
}

extern "C" tm_bico *bico_create(int64_t dim, int64_t n, int64_t k, int64_t nrandproj, int64_t coresetsize, int seed) {
  (void)nrandproj;
  tm_bico *b = new tm_bico();
  b->dim = dim; b->k = k; b->coreset = coresetsize; b->seed = (uint64_t)(uint32_t)seed;
  if (n > 0 && dim > 0) { b->rows.reserve((size_t)n * dim); b->weights.reserve((size_t)n); }
  return b;
}
extern "C" void bico_destroy(tm_bico *b) { delete b; }
extern "C" void bico_set_num_threads(int n) { (void)n; }
extern "C" void bico_set_rebuild_properties(tm_bico *b, uint32_t interval, double initial, double grow) { (void)b; (void)interval; (void)initial; (void)grow; }
extern "C" void bico_insert_line(tm_bico *b, const double *row, double weight) {
  if (!b || !row) return;
  b->rows.insert(b->rows.end(), row, row + b->dim);
  b->weights.push_back(weight);
}
// The BICO stand-in as one batched call: <= k weighted summary points of n weighted rows (host or device arrays in; the
// summary always comes back to HOST arrays: the host reads the count to size its next step, tilingencoder.pas:4168).
// n <= k: identity.  Else unweighted k-means++ seeding, at most max_iter weighted Lloyd updates, empty clusters dropped
// (oracle/tm_oracle.c: tmo_coreset_weighted defines the result bit for bit).
extern "C" int tm_coreset_weighted(const double *x, const double *w, int64_t n, int dim, int64_t k, int max_iter, uint64_t seed,
                                   double *centroids, double *weights, int64_t *count) {
  RC(require_gpu());
  if (!x || !w || !centroids || !count || n < 1 || n > 0x7fffffff || dim < 1 || dim > 1024 || k < 1 || k > 0x7fffffff || max_iter < 1 ||
      is_device_ptr(centroids) || is_device_ptr(weights))
    return fail(TM_ERR_ARG, "tm_coreset_weighted: bad argument (centroids / weights are host arrays)");
  std::lock_guard<std::recursive_mutex> lk(g_mu);
  Stage s(t_stream);
  if (n <= k) {   // every point is its own summary
    CU(cudaMemcpyAsync(centroids, x, (size_t)n * dim * 8, cudaMemcpyDefault, s.st));
    if (weights) CU(cudaMemcpyAsync(weights, w, (size_t)n * 8, cudaMemcpyDefault, s.st));
    CU(cudaStreamSynchronize(s.st));
    *count = n;
    return TM_OK;
  }
  const double *d_x = s.in(x, (size_t)n * dim), *d_w = s.in(w, (size_t)n);
  int32_t *d_labels = (int32_t *)s.temp((size_t)n * 4);
  std::vector<double> h_cent((size_t)k * dim), h_ws((size_t)k);
  double *d_cent = s.out(h_cent.data(), (size_t)k * dim), *d_ws = s.out(h_ws.data(), (size_t)k);
  if (s.err == TM_OK) s.err = kmeans_fit_dev(d_x, d_w, n, dim, (int)k, max_iter, nullptr, seed, 0, d_labels, d_cent, d_ws, nullptr, nullptr, s);
  RC(s.finish(true));
  int64_t m = 0;
  for (int64_t c = 0; c < k; ++c) {
    if (!(h_ws[c] > 0.0)) continue;   // drop empty clusters: the host accepts any count <= coresetsize (:4168)
    memcpy(centroids + (size_t)m * dim, &h_cent[(size_t)c * dim], (size_t)dim * sizeof(double));
    if (weights) weights[m] = h_ws[c];
    ++m;
  }
  *count = m;
  return TM_OK;
}

extern "C" int64_t bico_get_results(tm_bico *b, double *centroids, double *weights) {
  if (!b || !centroids) return 0;
  const int64_t n = (int64_t)b->weights.size();
  if (n == 0) return 0;
  int64_t m = 0;
  // bounded effort (8 weighted Lloyd updates): a coreset is a summary, not a converged clustering
  if (tm_coreset_weighted(b->rows.data(), b->weights.data(), n, (int)b->dim, b->coreset, 8, b->seed, centroids, weights, &m) != TM_OK) return 0;
  return m;
}
