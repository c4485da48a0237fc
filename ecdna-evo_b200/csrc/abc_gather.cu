// abc_gather.cu -- the one exchange step of the path (abc.md:57-78): the accepted prior draws of every
// GPU, packed on the device into fixed-stride records and all-gathered over NCCL (NVLink / NVSwitch).
//
// Packing is a count / scan / scatter compaction in index order (deterministic) followed by one warp per
// accepted draw copying its record: (index, rates, distances, summary statistics, final distribution).
// The collective is two ncclAllGather calls in one group on the context's stream - the per-rank counts and
// the per-rank record blocks of a fixed capacity - so no host synchronisation sits between the simulation,
// the packing and the exchange.  NCCL is loaded with dlopen at communicator creation: the library has no
// link-time dependency on it and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; the functions are resolved at run time

#include <cstring>
#include <mutex>
#include <string>

#include "engine.cuh"

using namespace ecdna;

namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
};

void nccl_load(NcclApi& api) {
  // a copy already loaded into the process (torch ships one) wins over the system library
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (api.handle) break;
  }
  for (const char* n : names) {
    if (api.handle) break;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.handle) {
    api.err = std::string("libnccl.so.2 not found: ") + dlerror();
    return;
  }
#define LOAD(field, sym)                                                    \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
  if (!api.field) api.err = std::string("missing NCCL symbol ") + sym;
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(AllGather, "ncclAllGather")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
}

NcclApi* nccl_api() {  // (loaded once, whichever thread asks first)
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] { nccl_load(api); });
  return &api;
}

#define NC(call)                                                                                           \
  do {                                                                                                     \
    ncclResult_t r__ = (call);                                                                             \
    if (r__ != ncclSuccess)                                                                                \
      return fail(ctx, ECDNA_B200_ERR_COMM, std::string(#call) + ": " + nccl_api()->GetErrorString(r__)); \
  } while (0)

constexpr int kScanBlock = 1024;

__global__ void flag_count(const uint8_t* flag, uint32_t n, uint32_t* block_counts) {
  const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  const int c = __syncthreads_count(i < n && flag[i] != 0);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)c;
}

// single block: exclusive scan of the block counts in place; the total goes to *total
__global__ void block_scan(uint32_t* block_counts, uint32_t n_blocks, uint32_t* total) {
  __shared__ uint32_t carry;
  __shared__ uint32_t wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_counts[i] : 0u;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if ((int)lane >= o) inc += u;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      const uint32_t s = lane < (blockDim.x >> 5) ? wsum[lane] : 0u;
      uint32_t si = s;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, si, o);
        if ((int)lane >= o) si += u;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    const uint32_t excl = carry + wsum[w] + inc - v;
    if (i < n_blocks) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void flag_scatter(const uint8_t* flag, uint32_t n, const uint32_t* block_offsets, uint32_t* out) {
  const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  const bool f = i < n && flag[i] != 0;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, f);
  __shared__ uint32_t wcount[32];
  if (lane == 0) wcount[w] = __popc(b);
  __syncthreads();
  uint32_t off = block_offsets[blockIdx.x];
  for (uint32_t j = 0; j < w; ++j) off += wcount[j];
  if (f) out[off + __popc(b & ((1u << lane) - 1u))] = i;
}

struct PackArgs {
  const uint32_t* idx;     // accepted run indices, ascending
  const uint32_t* count;   // how many
  uint32_t cap, rec_bins, hist_stride;
  uint64_t idx_begin;
  const float* rates;      // [n][4] or NULL (then the base rates)
  float base[4];
  ecdna_b200_results_t r;
  uint32_t* out;           // [cap][ECDNA_B200_ABC_REC_HEADER + rec_bins]
};

// one warp per accepted draw: the header by lane 0, the distribution by all lanes
__global__ void pack_records(const PackArgs a) {
  const uint32_t n = min(*a.count, a.cap);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = ECDNA_B200_ABC_REC_HEADER + a.rec_bins;
  for (uint32_t rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < n; rec += gridDim.x * (blockDim.x >> 5)) {
    const uint32_t run = a.idx[rec];
    uint32_t* o = a.out + (size_t)rec * stride;
    if (lane == 0) {
      const uint64_t idx = a.idx_begin + run;
      o[0] = (uint32_t)idx;
      o[1] = (uint32_t)(idx >> 32);
      for (int i = 0; i < 4; ++i) o[2 + i] = __float_as_uint(a.rates ? a.rates[(size_t)run * 4 + i] : a.base[i]);
      for (int i = 0; i < 4; ++i) o[6 + i] = a.r.abc_distance ? __float_as_uint(a.r.abc_distance[(size_t)run * 4 + i]) : 0u;
      o[10] = a.r.mean ? __float_as_uint(a.r.mean[run]) : 0u;
      o[11] = a.r.frequency ? __float_as_uint(a.r.frequency[run]) : 0u;
      o[12] = a.r.entropy ? __float_as_uint(a.r.entropy[run]) : 0u;
      o[13] = (a.r.nminus && a.r.nplus) ? (uint32_t)(a.r.nminus[run] + a.r.nplus[run]) : 0u;
      o[14] = a.r.kmax ? a.r.kmax[run] : 0u;
      o[15] = a.r.stop_reason ? a.r.stop_reason[run] : 0u;
    }
    const uint32_t* h = a.r.hist ? a.r.hist + (size_t)run * a.hist_stride : nullptr;
    for (uint32_t k = lane; k < a.rec_bins; k += 32u) o[ECDNA_B200_ABC_REC_HEADER + k] = (h && k < a.hist_stride) ? h[k] : 0u;
  }
}

}  // namespace

struct ecdna_b200_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

extern "C" {

int ecdna_b200_comm_unique_id(uint8_t id[ECDNA_B200_COMM_ID_BYTES]) {
  NcclApi* api = nccl_api();
  if (!api->err.empty() || !id) return ECDNA_B200_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == ECDNA_B200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  if (api->GetUniqueId(&u) != ncclSuccess) return ECDNA_B200_ERR_COMM;
  std::memcpy(id, &u, sizeof u);
  return ECDNA_B200_OK;
}

int ecdna_b200_comm_init(ecdna_b200_ctx* ctx, const uint8_t id[ECDNA_B200_COMM_ID_BYTES], int rank, int world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, ECDNA_B200_ERR_BAD_PARAMS, "bad communicator request");
  NcclApi* api = nccl_api();
  if (!api->err.empty()) return fail(ctx, ECDNA_B200_ERR_COMM, api->err);
  CU(cudaSetDevice(ctx->device));
  if (ctx->comm) { api->CommDestroy(ctx->comm->comm); delete ctx->comm; ctx->comm = nullptr; }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ecdna_b200_comm* c = new ecdna_b200_comm();
  c->rank = rank;
  c->world = world;
  const ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(ctx, ECDNA_B200_ERR_COMM, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
  }
  ctx->comm = c;
  return ECDNA_B200_OK;
}

void ecdna_b200_comm_release(ecdna_b200_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  NcclApi* api = nccl_api();
  if (api->err.empty() && ctx->comm->comm) api->CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
}

int ecdna_b200_abc_pack(ecdna_b200_ctx* ctx, const ecdna_b200_results_t* results_dev, const float* rates_dev,
                        const float base_rates[4], uint64_t idx_begin, uint64_t n_runs, uint32_t hist_stride,
                        uint32_t rec_bins, uint32_t capacity, uint32_t* records_dev, uint32_t* count_dev,
                        void* cuda_stream) {
  if (!ctx || !results_dev || !results_dev->abc_accept || !records_dev || !count_dev || n_runs == 0 ||
      n_runs >= (1ull << 32) || capacity == 0)
    return fail(ctx, ECDNA_B200_ERR_BAD_PARAMS, "bad pack request (abc_accept, records and count are required)");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  const uint32_t n = (uint32_t)n_runs;
  const uint32_t nb = (n + kScanBlock - 1) / kScanBlock;
  CU(ctx->pack_cnt.ensure((size_t)nb * 4));
  CU(ctx->pack_idx.ensure((size_t)n * 4));
  uint32_t* counts = (uint32_t*)ctx->pack_cnt.p;
  flag_count<<<nb, kScanBlock, 0, st>>>(results_dev->abc_accept, n, counts);
  block_scan<<<1, 1024, 0, st>>>(counts, nb, count_dev);
  flag_scatter<<<nb, kScanBlock, 0, st>>>(results_dev->abc_accept, n, counts, (uint32_t*)ctx->pack_idx.p);
  PackArgs a{};
  a.idx = (const uint32_t*)ctx->pack_idx.p;
  a.count = count_dev;
  a.cap = capacity;
  a.rec_bins = rec_bins;
  a.hist_stride = hist_stride;
  a.idx_begin = idx_begin;
  a.rates = rates_dev;
  for (int i = 0; i < 4; ++i) a.base[i] = base_rates ? base_rates[i] : 0.f;
  a.r = *results_dev;
  a.out = records_dev;
  const unsigned blocks = (unsigned)std::min<uint64_t>((uint64_t)ctx->sm_count * 8, ((uint64_t)std::min<uint64_t>(n, capacity) + 7) / 8);
  pack_records<<<blocks ? blocks : 1, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return ECDNA_B200_OK;
}

int ecdna_b200_abc_allgather(ecdna_b200_ctx* ctx, const uint32_t* records_dev, const uint32_t* count_dev, uint32_t rec_bins,
                             uint32_t capacity, uint32_t* all_records_dev, uint32_t* all_counts_dev, void* cuda_stream) {
  if (!ctx || !records_dev || !count_dev || !all_records_dev || !all_counts_dev || capacity == 0)
    return fail(ctx, ECDNA_B200_ERR_BAD_PARAMS, "bad all-gather request");
  if (!ctx->comm) return fail(ctx, ECDNA_B200_ERR_COMM, "no communicator: call ecdna_b200_comm_init first");
  NcclApi* api = nccl_api();
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
  const size_t bytes = (size_t)capacity * (ECDNA_B200_ABC_REC_HEADER + rec_bins) * 4;
  NC(api->GroupStart());
  NC(api->AllGather(count_dev, all_counts_dev, 1, ncclUint32, ctx->comm->comm, st));
  NC(api->AllGather(records_dev, all_records_dev, bytes, ncclUint8, ctx->comm->comm, st));
  NC(api->GroupEnd());
  return ECDNA_B200_OK;
}

}  // extern "C"
