// subsample.cuh -- EcDNADistribution::into_subsampled on the device (src/main.rs:110-123: after the final
// save, every --subsamples size gets a file with that many cells drawn from the final population).
//
// Native-mode definition (the oracle's orc_subsample restates it): n cells are drawn WITHOUT replacement,
// uniformly, from the final distribution (cells without ecDNA are class 0), i.e. a multivariate
// hypergeometric sample, built one cell at a time so that it is exact in integers: draw d picks
// u = bounded(N - d) and removes one cell of the first class k (natural order) whose cumulative count
// exceeds u.  When the sample is more than half of the population the cells NOT sampled are drawn instead.
// bounded(): Lemire's unbiased method on a 64-bit uniform from Philox4x32-10, key = seed, counter =
// (d, 0x40000000 + j + 65536 * (attempt / 2), run_lo, run_hi), words (0,1) then (2,3); j = index of the
// subsample size.  Slots >= 2^30 are never used by the event loop, so the streams do not overlap.
//
// One warp per (replicate, size); lane l owns a contiguous chunk of classes, the working counts live in
// the output row itself.
#pragma once
#include "ssa_kernel.cuh"

namespace ecdna {

struct SubArgs {
  uint32_t seed_lo, seed_hi;
  uint64_t idx_begin;
  uint32_t n_runs, n_sub, stride;
  const unsigned long long* sizes;  // [n_sub] device
  const uint32_t* hist;             // [n_runs][stride] final distributions
  uint32_t* out;                    // [n_runs][n_sub][stride]
};

__device__ __forceinline__ uint32_t sub_bounded(const SubArgs& a, uint32_t r0, uint32_t r1, uint32_t j, uint32_t d,
                                                uint32_t n) {
  uint32_t res = 0;
  for (uint32_t attempt = 0;; ++attempt) {
    const uint4 x = philox4x32_10(d, 0x40000000u + j + 65536u * (attempt >> 1), r0, r1, a.seed_lo, a.seed_hi);
    const unsigned long long v = (attempt & 1u) ? (((unsigned long long)x.z << 32) | x.w)
                                                : (((unsigned long long)x.x << 32) | x.y);
    const unsigned long long lo = v * (unsigned long long)n;
    res = (uint32_t)__umul64hi(v, (unsigned long long)n);
    if (lo >= n || attempt >= 25u) break;
    const unsigned long long t = (0ull - (unsigned long long)n) % n;
    if (lo >= t) break;
  }
  return res;
}

__global__ void __launch_bounds__(128) subsample_kernel(const SubArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const unsigned long long task = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= (unsigned long long)a.n_runs * a.n_sub) return;
  const uint32_t run = (uint32_t)(task / a.n_sub), j = (uint32_t)(task % a.n_sub);
  const uint32_t* h = a.hist + (size_t)run * a.stride;
  uint32_t* w = a.out + ((size_t)run * a.n_sub + j) * a.stride;
  const uint32_t chunk = (a.stride + 31u) / 32u;
  const uint32_t k_lo = min(lane * chunk, a.stride), k_hi = min(k_lo + chunk, a.stride);
  uint32_t ls = 0;
  for (uint32_t k = k_lo; k < k_hi; ++k) {
    const uint32_t v = __ldcg(h + k);  // written by the SSA kernel that ran just before on this stream
    w[k] = v;
    ls += v;
  }
  uint32_t P = ls;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(kFull, P, o);
    if ((int)lane >= o) P += up;
  }
  const uint32_t total = __shfl_sync(kFull, P, 31);
  const unsigned long long want = a.sizes[j];
  if (want >= total) return;  // the whole population (into_subsampled of a size >= the population)
  const uint32_t m = (uint32_t)want;
  const bool remove = m > total - m;
  const uint32_t draws = remove ? total - m : m;
  const uint64_t idx = a.idx_begin + run;
  const uint32_t r0 = (uint32_t)idx, r1 = (uint32_t)(idx >> 32);
  for (uint32_t d = 0; d < draws; ++d) {
    const uint32_t u = sub_bounded(a, r0, r1, j, d, total - d);
    const int lstar = __ffs(__ballot_sync(kFull, u < P)) - 1;
    if ((int)lane == lstar) {
      uint32_t rank = u - (P - ls);
      for (uint32_t k = k_lo; k < k_hi; ++k) {
        const uint32_t c = w[k];
        if (rank < c) { w[k] = c - 1u; break; }
        rank -= c;
      }
      ls -= 1u;
    }
    if ((int)lane >= lstar) P -= 1u;
  }
  if (!remove)
    for (uint32_t k = k_lo; k < k_hi; ++k) w[k] = __ldcg(h + k) - w[k];
}

}  // namespace ecdna
