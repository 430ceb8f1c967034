// palette_host.cpp -- HOST-side colour ordering of the palettes: TTilingEncoder.OptimizePalettes (tilingencoder.pas:4246-4432)
// over the reference's Powell minimiser (powell.pas: Bracket :56-146, BrentHelper :148-260, Brent :262-274, LinesearchPowell
// :294-324, PowellMinimize :326-385).  This is code the FreePascal host KEEPS (it is not behind a DLL); it lives in libtm_gtm.so
// next to the other host-side pieces so that an encode can be completed and verified here.  CPU code, one std::thread per
// hardware thread over the palettes like the reference's DoParallelLocalProc (:4415).
//
// It only permutes the colours INSIDE each palette (the picture does not change; index order, dithering tie-breaks and the
// compressibility of the stream do).  The minimiser works on a continuous vector whose rank order is the permutation, so the
// result depends on every floating-point detail of the search: the port keeps powell.pas's arithmetic order, its tolerances
// (scale = xtol = ftol = 1.0, :4381) and the reference semantics of FreePascal dynamic arrays -- `direc[n-1] := direc1`
// makes both names ONE array (no copy), which later direction updates write through.
#include "../../include/tm_gtm.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

namespace {

using Vec = std::vector<double>;
using VecRef = std::shared_ptr<Vec>;   // a FreePascal dynamic array variable: assignment shares, Copy() copies
using Fn1 = std::function<double(double)>;
using FnN = std::function<double(const Vec &)>;

inline double sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

// powell.pas:56-146
void bracket(const Fn1 &f, double xa, double xb, double &oa, double &ob, double &oc) {
  const int MaxIter = 1000;
  const double GrowLimit = 110, Gold = (1 + std::sqrt(5.0)) / 2, Small = 1e-21;
  double fa = f(xa), fb = f(xb);
  if (fa < fb) { std::swap(xa, xb); std::swap(fa, fb); }
  double xc = xb + Gold * (xb - xa);
  double fc = f(xc);
  int iter = 0;
  while (fc < fb) {
    const double tmp1 = (xb - xa) * (fb - fc);
    const double tmp2 = (xb - xc) * (fb - fa);
    const double val = tmp2 - tmp1;
    const double denom = std::fabs(val) < Small ? 2 * Small : 2 * val;
    double w = xb - ((xb - xc) * tmp2 - (xb - xa) * tmp1) / denom;
    const double wlim = xb + GrowLimit * (xc - xb);
    if (iter > MaxIter) throw std::runtime_error("bracket: Too many iterations");
    ++iter;
    double fw = 0;
    if ((w - xc) * (xb - w) > 0) {
      fw = f(w);
      if (fw < fc) { xa = xb; xb = w; fa = fb; fb = fw; break; }
      else if (fw > fb) { xc = w; fc = fw; break; }
      w = xc + Gold * (xc - xb);
      fw = f(w);
    } else if ((w - wlim) * (wlim - xc) >= 0) {
      w = wlim;
      fw = f(w);
    } else if ((w - wlim) * (xc - w) > 0) {
      fw = f(w);
      if (fw < fc) {
        xb = xc; xc = w; w = xc + Gold * (xc - xb);
        fb = fc; fc = fw; fw = f(w);
      }
    } else {
      w = xc + Gold * (xc - xb);
      fw = f(w);
    }
    xa = xb; xb = xc; xc = w;
    fa = fb; fb = fc; fc = fw;
  }
  if (xa > xc) { std::swap(xa, xc); std::swap(fa, fc); }
  oa = xa; ob = xb; oc = xc;
}

// powell.pas:148-260 -> (x, fx)
void brent_helper(const Fn1 &f, double a, double x, double b, double fx, double xtol, int maxiter, double &ox, double &ofx) {
  const double CG = (3 - std::sqrt(5.0)) / 2;
  if (a > b) std::swap(a, b);
  double w = x, v = x, fw = fx, fv = fx, deltax = 0, rat = 0;
  int iter = 0;
  while (iter < maxiter) {
    const double xmid = 0.5 * (a + b);
    if (std::fabs(x - xmid) <= 2 * xtol - 0.5 * (b - a)) break;
    if (std::fabs(deltax) <= xtol) {
      deltax = x >= xmid ? a - x : b - x;
      rat = CG * deltax;
    } else {
      double tmp1 = (x - w) * (fx - fv);
      double tmp2 = (x - v) * (fx - fw);
      double p = (x - v) * tmp2 - (x - w) * tmp1;
      tmp2 = 2 * (tmp2 - tmp1);
      if (tmp2 > 0) p = -p;
      tmp2 = std::fabs(tmp2);
      const double dx_temp = deltax;
      deltax = rat;
      if (p > tmp2 * (a - x) && p < tmp2 * (b - x) && std::fabs(p) < std::fabs(0.5 * tmp2 * dx_temp)) {
        rat = p / tmp2;
        const double u = x + rat;
        if (u - a < xtol || b - u < xtol) rat = sign(xmid - x) * xtol;
      } else {
        deltax = x >= xmid ? a - x : b - x;
        rat = CG * deltax;
      }
    }
    const double u = std::fabs(rat) > xtol ? x + rat : x + sign(rat) * xtol;
    const double fu = f(u);
    if (fu > fx) {
      if (u < x) a = u; else b = u;
      if (fu <= fw || w == x) { v = w; w = u; fv = fw; fw = fu; }
      else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
    } else {
      if (u >= x) a = x; else b = x;
      v = w; w = x; x = u;
      fv = fw; fw = fx; fx = fu;
    }
    ++iter;
  }
  ox = x; ofx = fx;
}

// powell.pas:294-324: minimise along p + alpha xi; xi is scaled by the step and p moved, both IN PLACE
double linesearch_powell(const FnN &f, Vec &p, Vec &xi, double xtol) {
  const size_t n = p.size();
  double sos = 0;
  for (double v : xi) sos += v * v;                 // SumOfSquares
  const double sqsos = std::sqrt(sos);
  double atol = 1.0;
  if (sqsos != 0) atol = 5 * xtol / sqsos;
  atol = std::min(0.1, atol);
  Vec tmp(n);
  const Fn1 along = [&](double t) {
    for (size_t i = 0; i < n; ++i) tmp[i] = p[i] + t * xi[i];
    return f(tmp);
  };
  double a, b, c;
  bracket(along, 0, 1, a, b, c);                    // Brent (:262-274)
  const double fb = along(b);
  double alpha, fret;
  brent_helper(along, a, b, c, fb, atol, 100, alpha, fret);
  for (size_t i = 0; i < n; ++i) { xi[i] = xi[i] * alpha; p[i] = p[i] + xi[i]; }
  return fret;
}

// powell.pas:326-385 -> final function value; x is updated in place
double powell_minimize(const FnN &f, Vec &x, double scale, double xtol, double ftol, int maxiter) {
  const size_t n = x.size();
  VecRef direc1 = std::make_shared<Vec>(n, 0.0);
  Vec tmp(n, 0.0);
  std::vector<VecRef> direc(n);
  for (size_t i = 0; i < n; ++i) { direc[i] = std::make_shared<Vec>(n, 0.0); (*direc[i])[i] = scale; }
  double fval = f(x);
  Vec x1 = x;                                       // Copy(x)
  int iter = 0;
  for (;;) {
    const double fx = fval;
    size_t bigind = 0;
    double delta = 0;
    for (size_t i = 0; i < n; ++i) {
      const double fx2 = fval;
      fval = linesearch_powell(f, x, *direc[i], xtol);
      if (fx2 - fval > delta) { delta = fx2 - fval; bigind = i; }
    }
    ++iter;
    if (fx - fval <= ftol || iter >= maxiter) break;
    for (size_t i = 0; i < n; ++i) {
      (*direc1)[i] = x[i] - x1[i];
      tmp[i] = x[i] + (*direc1)[i];
      x1[i] = x[i];
    }
    const double fx2 = f(tmp);
    if (fx > fx2) {
      double t = 2 * (fx + fx2 - 2 * fval);
      double temp = fx - fval - delta;
      t = t * temp * temp;
      temp = fx - fx2;
      t = t - delta * temp * temp;
      if (t < 0) {
        fval = linesearch_powell(f, x, *direc1, xtol);
        direc[bigind] = direc[n - 1];               // shares the array, as the Pascal assignment does
        direc[n - 1] = direc1;
      }
    }
  }
  return fval;
}

inline double pascal_round(double v) { return std::nearbyint(v); }   // Round: half to even (default rounding mode)

}  // namespace

// palettes [n_pal][pal_size] int32 0x00BBGGRR (null colours as they are), reordered in place.  Returns the number of outer
// iterations (:4394-4429), -1 on bad arguments.
extern "C" int tmh_optimize_palettes(int32_t *palettes, int n_pal, int pal_size, int n_threads) {
  if (!palettes || n_pal < 1 || pal_size < 2 || pal_size > 64 * 64) return -1;   // PalR/G/B hold Sqr(cTileWidth) = 64 columns in the reference
  const size_t P = (size_t)n_pal, S = (size_t)pal_size;
  std::vector<int32_t> newpal(P * S);
  std::vector<double> fv(P, 0.0);
  // mean of all palette colours (:4396-4412): the sum over every colour of every palette, divided by PaletteSize
  uint64_t meanR = 0, meanG = 0, meanB = 0;
  for (size_t i = 0; i < P * S; ++i) {
    const uint32_t c = (uint32_t)palettes[i];
    meanR += c & 255; meanG += (c >> 8) & 255; meanB += (c >> 16) & 255;
  }
  meanR /= S; meanG /= S; meanB /= S;
  int hw = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  int iteration = 0;
  double prev_fsum = 0, fsum = 0;
  do {
    prev_fsum = std::max(fsum, prev_fsum);
    ++iteration;
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (;;) {
        const size_t a = next.fetch_add(1);
        if (a >= P) break;
        // accumulate the whole palette set except the palette that will be permuted (:4357-4377)
        std::vector<uint64_t> palR(S, 0), palG(S, 0), palB(S, 0);
        for (size_t p = 0; p < P; ++p) {
          if (p == a) continue;
          for (size_t c = 0; c < S; ++c) {
            const uint32_t col = (uint32_t)palettes[p * S + c];
            palR[c] += col & 255; palG[c] += (col >> 8) & 255; palB[c] += (col >> 16) & 255;
          }
        }
        struct CI { int index, count; };
        std::vector<CI> perm(S);
        // PowellOP (:4264-4305): rank of x -> permutation; maximise the accumulated per-column standard deviation
        const FnN op = [&](const Vec &x) {
          perm[0] = {0, 0};
          for (size_t c = 1; c < S; ++c) perm[c] = {(int)c, (int)pascal_round(x[c - 1] * 1000)};
          std::sort(perm.begin(), perm.end(), [](const CI &l, const CI &r) { return l.count != r.count ? l.count < r.count : l.index < r.index; });
          uint64_t sdR = 0, sdG = 0, sdB = 0;
          for (size_t c = 0; c < S; ++c) {
            const int32_t col = palettes[a * S + (size_t)perm[c].index];
            newpal[a * S + c] = col;
            const uint32_t u = (uint32_t)col;
            const uint64_t dr = palR[c] + (u & 255) - meanR, dg = palG[c] + ((u >> 8) & 255) - meanG, db = palB[c] + ((u >> 16) & 255) - meanB;
            sdR += dr * dr; sdG += dg * dg; sdB += db * db;      // UInt64 arithmetic: a negative difference squares to the same value mod 2^64
          }
          const double res = (299.0 * std::sqrt((double)sdR / (double)S) + 587.0 * std::sqrt((double)sdG / (double)S) +
                              114.0 * std::sqrt((double)sdB / (double)S)) / 1000.0;
          return -res;
        };
        Vec x(S - 1);
        for (size_t c = 1; c < S; ++c) x[c - 1] = (double)c;
        try {
          powell_minimize(op, x, 1.0, 1.0, 1.0, 0x7fffffff);
        } catch (const std::exception &) {
          // bracket's iteration limit (:78): the reference would abort the encode; keep the best point reached instead
        }
        fv[a] = -op(x);                                            // also leaves NewPal[a] = the palette ordered by the final x (:4383)
      }
    };
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)hw, P);
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    fsum = 0;
    for (size_t p = 0; p < P; ++p) fsum += fv[p];
    std::copy(newpal.begin(), newpal.end(), palettes);              // FPalettes := NewPal, also on the last (non-improving) pass
    fsum /= (double)P;
  } while (!(fsum <= prev_fsum));
  return iteration;
}
