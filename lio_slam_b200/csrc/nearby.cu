// nearby.cu — which keyframes form the local map: mapOptimization::extractNearby (mapOptmization.cpp:1519-1554)
// and the selection half of extractCloud (:1558-1565); SURVEY §8 row f4.  The reference rebuilds a KD-tree over
// all key poses every scan; on a long run (10^4 key poses) that and the per-centroid nearest searches are the
// host-side cost that remains once the registration takes half a millisecond.
//
//   nb_dist_kernel    thread per key pose: L2_Simple distance to the newest pose -> radix-sort key (f32 bits of
//                     d^2 for a hit, 0xffffffff otherwise: FLANN's radius set is dist^2 < (float)(r*r), sorted by
//                     distance; the stable sort keeps equal distances in index order), hit count, and the newest
//                     pose that fails the 10 s recency test (:1547)
//   radix sort        (sort.cu) -> hits in distance order -> nb_gather_kernel -> VoxelGrid (voxel.cu; a few
//                     hundred poses: the single-block kernel) = downSizeFilterSurroundingKeyPoses (:1535-1536)
//   nb_snap_kernel    block per centroid: nearest key pose over ALL poses (nearestKSearch(pt, 1), :1539-1540;
//                     ties: lower index), then extractCloud's distance guard (:1562)
//   nb_recent_kernel  the trailing poses younger than 10 s (:1544-1551), newest first, same guard
// The host receives the ordered id list (duplicates kept, as the reference concatenates them).
#include "common.cuh"

#include <math_constants.h>

namespace liogpu {

namespace {

struct NbState {
  unsigned n_hits;
  int first_recent;  // smallest i0 such that every pose i >= i0 passes the recency test
};

__device__ __forceinline__ float nb_d2(const float4 a, const float4 b) {  // FLANN L2_Simple
  float r = 0.f;
  float d = a.x - b.x; r += d * d;
  d = a.y - b.y;       r += d * d;
  d = a.z - b.z;       r += d * d;
  return r;
}
__device__ __forceinline__ float nb_point_distance(const float4 p1, const float4 p2) {  // lib/common_lib.cpp:34-37
  return (float)sqrt((double)((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y) + (p1.z - p2.z) * (p1.z - p2.z)));
}

__global__ void nb_init_kernel(NbState* st) {
  st->n_hits = 0;
  st->first_recent = 0;
}

__global__ void __launch_bounds__(256)
nb_dist_kernel(const float4* __restrict__ key3d, const double* __restrict__ key_time, int n, double time_cur, float r2,
               uint32_t* __restrict__ keys, NbState* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d2 = nb_d2(key3d[n - 1], key3d[i]);
  const bool hit = d2 < r2;
  keys[i] = hit ? __float_as_uint(d2) : 0xffffffffu;
  const unsigned hm = __ballot_sync(__activemask(), hit);
  if (hit && (threadIdx.x & 31) == (__ffs(hm) - 1)) atomicAdd(&st->n_hits, (unsigned)__popc(hm));
  if (!(time_cur - key_time[i] < 10.0)) atomicMax(&st->first_recent, i + 1);
}

__global__ void __launch_bounds__(256)
nb_gather_kernel(const float4* __restrict__ key3d, const uint32_t* __restrict__ perm, int n_hits, float4* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_hits) out[j] = key3d[perm[j]];
}

constexpr int SNAP_THREADS = 256;
__global__ void __launch_bounds__(SNAP_THREADS)
nb_snap_kernel(const float4* __restrict__ ds, const float4* __restrict__ key3d, int n, float radius, int* __restrict__ ids) {
  __shared__ unsigned long long sh[SNAP_THREADS / 32];
  const float4 pt = ds[blockIdx.x];
  unsigned long long best = ~0ULL;
  for (int i = threadIdx.x; i < n; i += SNAP_THREADS) {
    const unsigned long long k = ((unsigned long long)__float_as_uint(nb_d2(pt, key3d[i])) << 32) | (unsigned)i;
    best = k < best ? k : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
    best = u < best ? u : best;
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < SNAP_THREADS / 32; ++k) best = sh[k] < best ? sh[k] : best;
    const int idx = (int)(unsigned)(best & 0xffffffffULL);
    const bool keep = !(nb_point_distance(pt, key3d[n - 1]) > radius);  // :1562
    ids[blockIdx.x] = keep ? (int)key3d[idx].w : -1;                    // pt.intensity = that pose's intensity (:1541, 1565)
  }
}

__global__ void __launch_bounds__(256)
nb_recent_kernel(const float4* __restrict__ key3d, int n, float radius, const NbState* __restrict__ st, int* __restrict__ ids,
                 int* __restrict__ n_recent) {
  const int i0 = st->first_recent;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;  // k-th newest pose
  if (k == 0) *n_recent = n - i0;
  const int i = n - 1 - k;
  if (i < i0) return;
  const bool keep = !(nb_point_distance(key3d[i], key3d[n - 1]) > radius);
  ids[k] = keep ? (int)key3d[i].w : -1;
}

}  // namespace

// key3d: device, packed (x, y, z, intensity = keyframe index); h_times: host.  ids (host, cap entries) receives the
// keyframe indices in concatenation order; *n_ids the count (also when it exceeds cap -> LIOGPU_E_CAPACITY).
int extract_nearby_dev(Ctx* c, const float4* key3d, int n, const double* h_times, double time_cur, float radius,
                       float density, int* ids, int cap, int* n_ids) {
  *n_ids = 0;
  LIOGPU_CUDA_OK(c, c->keys0.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->keys1.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->vals0.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->vals1.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->lm_md.reserve((size_t)n * sizeof(double)));      // key times
  LIOGPU_CUDA_OK(c, c->lm_b.reserve((size_t)n * sizeof(float4)));       // hits in distance order
  LIOGPU_CUDA_OK(c, c->lm_flag.reserve((size_t)2 * n * sizeof(int) + 64));  // ids: centroids, then recents
  LIOGPU_CUDA_OK(c, c->lm_stats.reserve(4096));
  NbState* d_st = reinterpret_cast<NbState*>((char*)c->lm_stats.p + 512);
  int* d_nrecent = reinterpret_cast<int*>((char*)c->lm_stats.p + 640);
  double* d_times = c->lm_md.as<double>();
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(d_times, h_times, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
  const float r2 = (float)((double)radius * (double)radius);  // pcl::KdTreeFLANN::radiusSearch narrows radius^2 to f32
  nb_init_kernel<<<1, 1, 0, c->stream>>>(d_st);
  nb_dist_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(key3d, d_times, n, time_cur, r2, c->keys0.as<uint32_t>(), d_st);
  c->launches += 2;
  uint32_t *skeys = nullptr, *sperm = nullptr;
  LIOGPU_CUDA_OK(c, radix_sort_pairs(c, n, 32, nullptr, &skeys, &sperm));
  NbState* h_st = reinterpret_cast<NbState*>((char*)c->h_pinned + 28672);
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_st, d_st, sizeof(NbState), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  const int n_hits = (int)h_st->n_hits;
  const int n_recent = n - h_st->first_recent;
  nb_gather_kernel<<<div_up(n_hits > 0 ? n_hits : 1, 256), 256, 0, c->stream>>>(key3d, sperm, n_hits, c->lm_b.as<float4>());
  c->launches++;
  int n_ds = 0;
  bool overflow = false;
  int rc = voxel_downsample_dev(c, c->lm_b.as<float4>(), n_hits, density, c->lm_out, &n_ds, &overflow);
  if (rc) return rc;
  int* d_ids = c->lm_flag.as<int>();
  if (n_ds > 0) {
    nb_snap_kernel<<<n_ds, SNAP_THREADS, 0, c->stream>>>(c->lm_out.as<float4>(), key3d, n, radius, d_ids);
    c->launches++;
  }
  if (n_recent > 0) {
    nb_recent_kernel<<<div_up(n_recent, 256), 256, 0, c->stream>>>(key3d, n, radius, d_st, d_ids + n_ds, d_nrecent);
    c->launches++;
  }
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
  const int total = n_ds + n_recent;
  std::vector<int> tmp((size_t)(total > 0 ? total : 1));
  if (total > 0)
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(tmp.data(), d_ids, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
  int m = 0;
  for (int k = 0; k < total; ++k) {
    if (tmp[k] < 0) continue;
    if (m < cap && ids) ids[m] = tmp[k];
    ++m;
  }
  *n_ids = m;
  if (m > cap) { c->err = "liogpu_extract_nearby: id capacity too small"; return LIOGPU_E_CAPACITY; }
  return LIOGPU_OK;
}

}  // namespace liogpu
