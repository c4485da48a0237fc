// uniform_replay.cuh -- replay of the reference's OWN random stream on the reference's own state layout.
//
// rng_mode = ECDNA_B200_RNG_UNIFORMS: the caller hands over, per replicate, every u64 the reference's
// generator produced (what a recording `RngCore` wrapped around the `ChaCha8Rng` of src/main.rs:57-58
// sees), and this kernel re-runs the replicate exactly as the reference does: per-cell `u16` vector
// (memory.md:5-8) in HBM, uniform index + swap-remove (proliferation.rs:57), daughters pushed k1 then k2
// (proliferation.rs:85-88,109), one Exp(rate) per reaction in sosa's order, Binomial(2k, 1/2) by
// rand_distr's BINV/BTPE.  Because the per-cell order is kept, the raw stream maps to the same cells
// -- something a histogram cannot do.  It is a verification mode, one thread per
// replicate, not a throughput path.
//
// The conversions from u64 to variates restate rand 0.8.5 / rand_distr 0.4.3 from their published
// algorithms [RECALL R2, R5, R6 in SURVEY 8c]: parity with the real crates is unpinned; parity with the
// CPU oracle's reference-layout configuration is bit-exact (tests/test_gpu_parity.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ssa_kernel.cuh"

namespace ecdna {

struct UrArgs {
  float rate[4];
  const float* rates_per_run;
  uint32_t segregation, cells_stop, max_iter_m1;
  float max_time;
  uint32_t n_runs;
  uint32_t n_init;
  const uint32_t* init_k;
  const uint32_t* init_c;
  uint32_t init_nminus;
  const uint64_t* stream;      // all runs' u64, back to back
  const uint64_t* stream_off;  // [n_runs + 1]
  uint16_t* cells;             // [n_runs][cap] per-cell copy numbers
  uint64_t cap;
  const double* zig_x;  // 257-entry ziggurat tables of Exp1 (layer edges and pdf values)
  const double* zig_f;
  uint32_t hist_stride;
  unsigned long long* totals;
  ecdna_b200_results_t out;
};

struct U64Stream {
  const uint64_t* p;
  uint64_t len, pos;
  bool dry;
  __device__ uint64_t next() {
    if (pos >= len) { dry = true; return 0x8000000000000000ull; }
    return p[pos++];
  }
  // rand 0.8.5 Standard f64: 53 random bits scaled by 2^-53
  __device__ double f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  // rand 0.8.5 UniformInt<usize>::sample_single: widening multiply, conservative zone
  __device__ uint64_t below(uint64_t n) {
    const uint64_t zone = (n << __clzll((long long)n)) - 1;
    for (;;) {
      const uint64_t v = next();
      if (dry) return 0;
      if (v * n <= zone) return __umul64hi(v, n);
    }
  }
  // rand 0.8.5 UniformFloat<f64>::sample: 52 mantissa bits -> [1,2) - 1, then * scale + low
  __device__ double uniform(double low, double scale) {
    const double v12 = __longlong_as_double((long long)((next() >> 12) | 0x3FF0000000000000ull));
    return __dadd_rn(__dmul_rn(v12 - 1.0, scale), low);
  }
};

// rand_distr 0.4.3 Exp1: 256-layer ziggurat
__device__ inline double ur_exp1(U64Stream& g, const double* zx, const double* zf) {
  for (;;) {
    const uint64_t bits = g.next();
    if (g.dry) return 0.0;
    const int i = (int)(bits & 0xff);
    const double u12 = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull));
    const double u = u12 - (1.0 - 2.220446049250313e-16 / 2.0);
    const double xx = __dmul_rn(u, zx[i]);
    if (xx < zx[i + 1]) return xx;
    if (i == 0) return 7.69711747013104972 - log(g.f64());
    if (__dadd_rn(zf[i + 1], __dmul_rn(zf[i] - zf[i + 1], g.f64())) < exp(-xx)) return xx;
  }
}

__device__ inline double ur_stirling(double a) {
  const double a2 = a * a;
  return (13860. - (462. - (132. - (99. - 140. / a2) / a2) / a2) / a2) / a / 166320.;
}

// rand_distr 0.4.3 Binomial::sample for p = 1/2: BINV below n*p < 10, BTPE otherwise
__device__ inline uint64_t ur_binomial_half(U64Stream& g, uint64_t n_int) {
  const double p = 0.5, q = 0.5;
  const double n = (double)n_int;
  if (n * p < 10.0) {
    const double s = p / q;
    const double a = (double)(n_int + 1) * s;
    double r = scalbn(1.0, -(int)n_int);  // q^n with q = 1/2, exact like powi
    double u = g.f64();
    long long xv = 0;
    while (u > r) {
      u -= r;
      xv += 1;
      r *= a / (double)xv - s;
    }
    return (uint64_t)xv;
  }
  const double np = n * p, npq = np * q, fm = np + p;
  const long long m = (long long)fm;
  const double p1 = floor(2.195 * sqrt(npq) - 4.6 * q) + 0.5;
  const double xm = (double)m + 0.5, xl = xm - p1, xr = xm + p1;
  const double c = 0.134 + 20.5 / (15.3 + (double)m);
  const double p2 = p1 * (1.0 + 2.0 * c);
  const double al = (fm - xl) / (fm - xl * p), ar = (xr - fm) / (xr * q);
  const double ll = al * (1.0 + 0.5 * al), lr = ar * (1.0 + 0.5 * ar);
  const double p3 = p2 + c / ll;
  const double p4 = p3 + c / lr;
  long long y;
  for (;;) {
    const double u = g.uniform(0.0, p4);
    double v = g.uniform(0.0, 1.0);
    if (g.dry) return 0;
    if (!(u > p1)) { y = (long long)(xm - p1 * v + u); break; }
    if (!(u > p2)) {
      const double xx = xl + (u - p1) / c;
      v = v * c + 1.0 - fabs(xx - xm) / p1;
      if (v > 1.0) continue;
      y = (long long)xx;
    } else if (!(u > p3)) {
      y = (long long)(xl + log(v) / ll);
      if (y < 0) continue;
      v *= (u - p2) * ll;
    } else {
      y = (long long)(xr - log(v) / lr);
      if (y > 0 && (uint64_t)y > n_int) continue;
      v *= (u - p3) * lr;
    }
    const long long kk = y > m ? y - m : m - y;
    if (!(kk > 20 && (double)kk < 0.5 * npq - 1.0)) {
      const double s = p / q, a = s * (n + 1.0);
      double ff = 1.0;
      if (m < y) { for (long long i = m + 1; i <= y; ++i) ff *= a / (double)i - s; }
      else if (m > y) { for (long long i = y + 1; i <= m; ++i) ff /= a / (double)i - s; }
      if (v > ff) continue;
      break;
    }
    const double kd = (double)kk;
    const double rho = (kd / npq) * ((kd * (kd / 3.0 + 0.625) + 1.0 / 6.0) / npq + 0.5);
    const double t = -0.5 * kd * kd / npq;
    const double alpha = log(v);
    if (alpha < t - rho) break;
    if (alpha > t + rho) continue;
    const double x1 = (double)(y + 1), f1 = (double)(m + 1);
    const double zz = (double)((long long)n + 1 - m), w = (double)((long long)n - y + 1);
    const double bound = xm * log(f1 / x1) + (n - (double)m + 0.5) * log(zz / w) +
                         (double)(y - m) * log(w * p / (x1 * q)) + ur_stirling(f1) + ur_stirling(zz) -
                         ur_stirling(x1) - ur_stirling(w);
    if (alpha > bound) continue;
    break;
  }
  return (uint64_t)y;
}

__global__ void __launch_bounds__(64) uniform_replay_kernel(const __grid_constant__ UrArgs a) {
  const uint32_t run = blockIdx.x * blockDim.x + threadIdx.x;
  if (run >= a.n_runs) return;
  U64Stream g;
  g.p = a.stream + a.stream_off[run];
  g.len = a.stream_off[run + 1] - a.stream_off[run];
  g.pos = 0;
  g.dry = false;
  uint16_t* cells = a.cells + (size_t)run * a.cap;
  float rate[4];
  for (int i = 0; i < 4; ++i) rate[i] = a.rates_per_run ? a.rates_per_run[(size_t)run * 4 + i] : a.rate[i];

  // EcDNADistribution::new expands the histogram into the per-cell vector, in the order given
  uint64_t nminus = a.init_nminus, nplus = 0, hash = 0, chain = 0, sum_k = 0;
  uint32_t kmax = 0, n_div = 0, n_death = 0, iter = 0;
  for (uint32_t i = 0; i < a.n_init; ++i) {
    const uint32_t k = a.init_k[i], c = a.init_c[i];
    for (uint32_t j = 0; j < c && nplus < a.cap; ++j) cells[nplus++] = (uint16_t)k;
    hash += hist_weight(k) * c;
    kmax = max(kmax, k);
  }
  float time = 0.f;
  uint32_t stop;
  for (;;) {
    const uint64_t n_cells = nminus + nplus;
    if (n_cells == 0) { stop = ECDNA_B200_STOP_NO_INDIVIDUALS; break; }
    if (iter >= a.max_iter_m1) { stop = ECDNA_B200_STOP_MAX_ITERS; break; }
    if (time >= a.max_time) { stop = ECDNA_B200_STOP_MAX_TIME; break; }
    if (n_cells >= a.cells_stop) { stop = ECDNA_B200_STOP_MAX_CELLS; break; }
    // sosa: one waiting time per reaction, in the order of main.rs:140-145; first minimum wins
    float best = __uint_as_float(kInfBits);
    uint32_t evt = 0xFFFFFFFFu;
    for (int i = 0; i < 4; ++i) {
      const float lam = __fmul_rn(rate[i], (i & 1) ? __ull2float_rn(nplus) : __ull2float_rn(nminus));
      const uint32_t lb = __float_as_uint(lam), ex = (lb >> 23) & 0xFFu;
      float t;
      if (ex != 0u && ex != 255u) t = __fmul_rn(__double2float_rn(ur_exp1(g, a.zig_x, a.zig_f)), __fdiv_rn(1.0f, lam));
      else t = lb == kInfBits ? 0.f : __uint_as_float(kInfBits);
      if (t < best) { best = t; evt = (uint32_t)i; }
    }
    if (g.dry) { stop = ECDNA_B200_STOP_REPLAY_END; break; }
    if (evt == 0xFFFFFFFFu) { stop = ECDNA_B200_STOP_ABSORBING; break; }
    if (evt == ECDNA_B200_EV_BIRTH_NMINUS) {
      nminus += 1;
    } else if (evt == ECDNA_B200_EV_DEATH_NMINUS) {
      nminus -= 1;
    } else {
      // (the per-cell arena holds max_cells + 2 cells and the run stops at max_cells: checked before
      //  anything is touched, so an overflow reports the pre-event state)
      if (nplus + 1 > a.cap) { stop = ECDNA_B200_STOP_HIST_OVERFLOW; break; }
      const uint64_t idx = g.below(nplus);
      if (g.dry) { stop = ECDNA_B200_STOP_REPLAY_END; break; }
      sum_k += (uint64_t)kmax + 1u;
      const uint32_t k = cells[idx];  // pick_remove_random_nplus: swap_remove
      cells[idx] = cells[nplus - 1];
      nplus -= 1;
      hash -= hist_weight(k);
      if (evt == ECDNA_B200_EV_DEATH_NPLUS) {
        n_death += 1;
      } else {
        n_div += 1;
        // (the reference panics here, proliferation.rs:63-67, with the cell already taken out at :57; the
        //  histogram kernel and the oracle report the same state)
        if (k >= 32768u) { stop = ECDNA_B200_STOP_COPY_OVERFLOW; break; }
        const uint32_t n = 2u * k;
        uint32_t k1, k2;
        bool uneven;
        if (a.segregation == ECDNA_B200_SEG_DETERMINISTIC) {
          k1 = k2 = k;
          uneven = false;
        } else {
          for (;;) {
            k1 = (uint32_t)ur_binomial_half(g, n);
            k2 = n - k1;
            uneven = (k1 == 0 || k2 == 0);
            if (g.dry || !(uneven && a.segregation == ECDNA_B200_SEG_BINOMIAL_NO_UNEVEN)) break;
          }
          if (g.dry) { stop = ECDNA_B200_STOP_REPLAY_END; break; }
        }
        if (!uneven) {
          cells[nplus++] = (uint16_t)k1;
          cells[nplus++] = (uint16_t)k2;
          hash += hist_weight(k1) + hist_weight(k2);
          kmax = max(kmax, max(k1, k2));
        } else {
          if (a.segregation == ECDNA_B200_SEG_BINOMIAL) nminus += 1;
          cells[nplus++] = (uint16_t)n;
          hash += hist_weight(n);
          kmax = max(kmax, n);
        }
      }
    }
    time = __fadd_rn(time, best);
    chain = chain_step(chain, hash, (uint32_t)nminus, time);
    iter += 1;
  }
  const ecdna_b200_results_t& o = a.out;
  uint32_t flags = 0;
  if (o.hist) {
    uint32_t* h = o.hist + (size_t)run * a.hist_stride;
    for (uint32_t k = 0; k < a.hist_stride; ++k) h[k] = 0;
    if (a.hist_stride) h[0] = (uint32_t)nminus;
    for (uint64_t i = 0; i < nplus; ++i) {
      const uint32_t k = cells[i];
      if (k < a.hist_stride) h[k] += 1;
    }
  }
  if (kmax >= a.hist_stride) flags |= ECDNA_B200_FLAG_HIST_TRUNCATED;
  if (o.stop_reason) o.stop_reason[run] = stop | flags;
  if (o.nminus) o.nminus[run] = nminus;
  if (o.nplus) o.nplus[run] = nplus;
  if (o.time) o.time[run] = time;
  if (o.n_events) o.n_events[run] = iter;
  if (o.kmax) o.kmax[run] = kmax;
  if (o.hash) o.hash[run] = hash;
  if (o.chain) o.chain[run] = chain;
  if (o.sum_k) o.sum_k[run] = sum_k;
  if (o.n_div) o.n_div[run] = n_div;
  if (o.n_death) o.n_death[run] = n_death;
  atomicAdd(a.totals + 0, (unsigned long long)iter);
  atomicAdd(a.totals + 1, (unsigned long long)sum_k);
  atomicAdd(a.totals + 2, (unsigned long long)n_div);
  atomicAdd(a.totals + 3, (unsigned long long)n_death);
  atomicAdd(a.totals + 7, 1ull);
}

}  // namespace ecdna
