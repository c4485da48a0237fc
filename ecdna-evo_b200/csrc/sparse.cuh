// sparse.cuh -- the sparse return of the distributions (include/ecdna_b200.h, ecdna_b200_run_sparse).
//
// The dense result columns hist [n][stride], snap_hist [n][S][stride] and sub_hist [n][U][stride] stay on the
// device; these kernels measure every distribution (first / last occupied copy number, cells), lay the windows
// out back to back (an exclusive prefix sum over the window lengths) and pack them into one arena, so that one
// 40-byte descriptor and the occupied bins per distribution cross the bus instead of `stride` words: what
// `save` (reference src/process.rs:31-55) writes is a map over the occupied copy numbers only.
//
// Rows are numbered finals first, [0, n), then the snapshots [n, n + n S), then the samples.  All three passes
// stream through HBM once: one warp per row, coalesced 128-byte reads.
#pragma once
#include <cstdint>

#include "../../include/ecdna_b200.h"

namespace ecdna {

struct SparseArgs {
  uint32_t n_runs, n_snap, n_sub, stride;
  unsigned long long rows;
  const uint32_t* hist;
  const uint32_t* snap_hist;
  const uint32_t* sub_hist;
  const uint32_t* snap_count;
  const float* time;
  const float* snap_time;
  const uint32_t* stop;
  ecdna_b200_dist_t* desc;     // [rows]
  uint32_t* len;               // [rows] window lengths (the scan's input)
  unsigned long long* bsum;    // [blocks + 1] per-block sums, then their exclusive scan; [blocks] = total words
  uint32_t* arena;
};

constexpr int kSparseChunk = 2048;  // rows per block of the scan passes (256 threads x 8)

// which dense row a row number means, and what goes with it
struct SparseRow {
  const uint32_t* bins;
  float time;
  uint32_t run;
  bool taken;
};
__device__ __forceinline__ SparseRow sparse_row(const SparseArgs& a, unsigned long long r) {
  SparseRow o;
  if (r < a.n_runs) {
    o.run = (uint32_t)r;
    o.bins = a.hist + (size_t)r * a.stride;
    o.time = a.time[o.run];
    o.taken = true;
    return o;
  }
  r -= a.n_runs;
  const unsigned long long n_snap_rows = (unsigned long long)a.n_runs * a.n_snap;
  if (r < n_snap_rows) {
    o.run = (uint32_t)(r / a.n_snap);
    const uint32_t s = (uint32_t)(r % a.n_snap);
    o.bins = a.snap_hist + (size_t)r * a.stride;
    o.taken = s < a.snap_count[o.run];
    o.time = o.taken ? a.snap_time[r] : 0.f;
    return o;
  }
  r -= n_snap_rows;
  o.run = (uint32_t)(r / a.n_sub);
  o.bins = a.sub_hist + (size_t)r * a.stride;
  o.time = a.time[o.run];
  o.taken = true;
  return o;
}

// pass 1: one warp per distribution
__global__ void __launch_bounds__(256) sparse_measure(const SparseArgs a) {
  const unsigned long long r = (unsigned long long)blockIdx.x * 8u + (threadIdx.x >> 5);
  if (r >= a.rows) return;
  const uint32_t lane = threadIdx.x & 31u;
  const SparseRow row = sparse_row(a, r);
  uint32_t lo = 0xFFFFFFFFu, hi = 0;
  unsigned long long cells = 0;
  if (row.taken)
    for (uint32_t k = lane; k < a.stride; k += 32u) {
      const uint32_t c = row.bins[k];
      cells += c;
      if (c != 0 && k != 0) {
        lo = min(lo, k);
        hi = k;  // (k ascends within a lane)
      }
    }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, d));
    cells += __shfl_xor_sync(0xFFFFFFFFu, cells, d);
  }
  if (lane != 0) return;
  ecdna_b200_dist_t d;
  d.cells = cells;
  d.nminus = row.taken ? row.bins[0] : 0;
  d.offset = 0;
  d.time = row.time;
  d.k_len = hi ? hi - lo + 1u : 0u;
  d.k_min = hi ? (uint16_t)lo : (uint16_t)0;
  d.flags = (uint16_t)((row.taken ? ECDNA_B200_DIST_TAKEN : 0u) |
                       ((row.taken && (a.stop[row.run] & ECDNA_B200_FLAG_HIST_TRUNCATED)) ? ECDNA_B200_DIST_TRUNCATED : 0u));
  d.reserved = 0;
  a.desc[r] = d;
  a.len[r] = d.k_len;
}

// block-wide sum / exclusive scan of one value per thread (256 threads)
__device__ __forceinline__ unsigned long long block_excl_scan_256(unsigned long long v, unsigned long long* total) {
  __shared__ unsigned long long warp_sums[8];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= (uint32_t)d) inc += t;
  }
  if (lane == 31u) warp_sums[w] = inc;
  __syncthreads();
  unsigned long long base = 0, all = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if ((uint32_t)i < w) base += warp_sums[i];
    all += warp_sums[i];
  }
  __syncthreads();
  *total = all;
  return base + inc - v;
}

// pass 2a: words per chunk of kSparseChunk rows
__global__ void __launch_bounds__(256) sparse_chunk_sums(const SparseArgs a) {
  const unsigned long long first = (unsigned long long)blockIdx.x * kSparseChunk + (unsigned long long)threadIdx.x * 8u;
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (first + i < a.rows) s += a.len[first + i];
  unsigned long long total;
  block_excl_scan_256(s, &total);
  if (threadIdx.x == 0) a.bsum[blockIdx.x] = total;
}

// pass 2b: exclusive scan of the chunk sums in place, one block; bsum[n_chunks] = all words
__global__ void __launch_bounds__(256) sparse_scan_chunks(unsigned long long* bsum, uint32_t n_chunks) {
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_chunks; base += 256u) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < n_chunks ? bsum[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_excl_scan_256(v, &total);
    const unsigned long long carry = carry_s;
    if (i < n_chunks) bsum[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[n_chunks] = carry_s;
}

// pass 2c: every row's offset
__global__ void __launch_bounds__(256) sparse_offsets(const SparseArgs a) {
  const unsigned long long first = (unsigned long long)blockIdx.x * kSparseChunk + (unsigned long long)threadIdx.x * 8u;
  uint32_t l[8];
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    l[i] = first + i < a.rows ? a.len[first + i] : 0u;
    s += l[i];
  }
  unsigned long long total;
  unsigned long long off = a.bsum[blockIdx.x] + block_excl_scan_256(s, &total);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (first + i < a.rows) a.desc[first + i].offset = off;
    off += l[i];
  }
}

// pass 3: one warp per distribution copies its window into the arena
__global__ void __launch_bounds__(256) sparse_pack(const SparseArgs a) {
  const unsigned long long r = (unsigned long long)blockIdx.x * 8u + (threadIdx.x >> 5);
  if (r >= a.rows) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t n = a.len[r];
  if (n == 0) return;
  const ecdna_b200_dist_t d = a.desc[r];
  const SparseRow row = sparse_row(a, r);
  const uint32_t* src = row.bins + d.k_min;
  uint32_t* dst = a.arena + d.offset;
  for (uint32_t i = lane; i < n; i += 32u) dst[i] = src[i];
}

}  // namespace ecdna
