// localmap.cu — mapOptimization::publishLocalMap (mapOptmization.cpp:2442-2541), which the
// reference runs after EVERY registration (MO:504):
//
//   globalMapCloud  = concat_i transformPointCloud(surfCloudKeyFrames[i], cloudKeyPoses6D[i])   (MO:2462-2466)
//   transformed     = pcl::transformPointCloud(globalMapCloud, Affine3f{ Rz(-yaw) | -R(-yaw) t })   (MO:2474-2489)
//   localMapCloud   = PassThrough y( PassThrough x( transformed ) )                               (MO:2502-2507)
//   [ tempCloud     = StatisticalOutlierRemoval(meanK, stddevThreshold) ]                         (MO:2510-2516)
//   [ tempCloud     = VoxelGrid(localMappingSurfLeafSize) ]                                       (MO:2517-2540)
//
// Device layout: the keyframes are already resident (packed float4).  One kernel does both transforms and the
// crop predicate per point (the intermediate clouds of the reference never exist), an exclusive scan + scatter
// compacts in input order (PassThrough and StatisticalOutlierRemoval both keep input order).  The outlier
// filter needs each point's mean distance to its meanK nearest neighbours inside the same cloud: a sorted
// uniform grid over the cropped cloud, searched shell by shell until the (meanK+1)-th distance is certified,
// thread-per-point for the first shells and warp-per-point for the isolated points that need many shells.
// Only the DISTANCES enter the result, so equidistant neighbours need no tie rule here.
#include "common.cuh"
#include "pose_math.cuh"

#include <math_constants.h>

#include "shell_search.cuh"

namespace liogpu {

namespace {

struct YawFrame {
  float m[12];  // row-major 3x4 of transformMatrix (MO:2486-2488)
  float xmin, xmax, ymin, ymax;
};

// ---- stage 1: keyframe transform + yaw-frame transform + crop predicate --------------------------------
__global__ void __launch_bounds__(256)
lmap_transform_crop_kernel(const float4* const* __restrict__ srcs, const int* __restrict__ offs, int k,
                           const float* __restrict__ T12, YawFrame yf, long long total, float4* __restrict__ out,
                           uint32_t* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = k;  // offs[lo] <= i < offs[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(offs + mid) <= i) lo = mid; else hi = mid;
  }
  const float4 src = srcs[lo][i - __ldg(offs + lo)];
  float T[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) T[q] = __ldg(T12 + 12 * lo + q);
  const float4 p = apply_T(T, src);  // transformPointCloud, MO:862-864
  // pcl::transformPointCloud with an Affine3f (PCL >= 1.10 detail::Transformer<float>::se3): per output
  // coordinate x*c0 + (y*c1 + (z*c2 + c3)), the association of its SSE form.
  float4 q;
  q.x = p.x * yf.m[0] + (p.y * yf.m[1] + (p.z * yf.m[2] + yf.m[3]));
  q.y = p.x * yf.m[4] + (p.y * yf.m[5] + (p.z * yf.m[6] + yf.m[7]));
  q.z = p.x * yf.m[8] + (p.y * yf.m[9] + (p.z * yf.m[10] + yf.m[11]));
  q.w = p.w;
  // pcl::PassThrough "x" in [xmin, xmax] then "y" in [ymin, ymax], negative = false (MO:296-302):
  // non-finite points are dropped, limits are inclusive.
  bool keep = isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
  if (q.x < yf.xmin || q.x > yf.xmax) keep = false;
  if (q.y < yf.ymin || q.y > yf.ymax) keep = false;
  out[i] = q;
  flag[i] = keep ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
lmap_compact_kernel(const float4* __restrict__ in, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                    int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) out[pos[i]] = in[i];
}

// ---- stage 2: mean distance to the meanK nearest neighbours (pcl::StatisticalOutlierRemoval, first pass) ---

// K smallest squared distances seen so far, ascending, in registers (static indexing only).
template <int K>
struct TopK {
  float a[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; ++j) a[j] = CUDART_INF_F;
  }
  __device__ __forceinline__ void push(float v, float /*index bits*/ = 0.f) {
    if (v < a[K - 1]) {  // a value equal to the current worst leaves the multiset of the K smallest unchanged
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float lo = fminf(a[j], v);
        v = fmaxf(a[j], v);
        a[j] = lo;
      }
    }
  }
  __device__ __forceinline__ float at(int idx) const {
    float r = a[K - 1];
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (j == idx) r = a[j];
    return r;
  }
};

// dist_sum / mean_k exactly as PCL: sqrt in f32, accumulation in f64 in ascending order, neighbour 0 (the
// query itself) skipped, narrowed to f32.
template <int K>
__device__ __forceinline__ float mean_distance(const TopK<K>& top, int mean_k) {
  double s = 0.0;
#pragma unroll
  for (int j = 1; j < K; ++j)
    if (j <= mean_k) s += (double)__fsqrt_rn(top.a[j]);
  return (float)(s / (double)mean_k);
}

constexpr int SOR_THREADS = 128;
constexpr int SOR_WARP_SHELLS = 6;  // warp-cooperative search gives up after shells 0..5 (>= 5 cell edges covered)

// thread per point, in grid-sorted order (neighbouring threads share cells); up to rmax shells
template <int K>
__global__ void __launch_bounds__(SOR_THREADS)
sor_knn_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ cs, const GridParams* __restrict__ gp,
               int rmax, int mean_k, float* __restrict__ md, uint32_t* __restrict__ left_list,
               uint32_t* __restrict__ left_count) {
  __shared__ GridParams g;
  if (threadIdx.x == 0) g = *gp;
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n_points) return;
  const float4 p = __ldg(sorted + j);
  const HomeCell hc = home_cell(g, p);
  TopK<K> top;
  top.init();
  bool done = false;
  for (int R = 0; R <= rmax && !done; ++R) {
    const int z0 = max(hc.cz - R, 0), z1 = min(hc.cz + R, g.nz - 1);
    const int y0 = max(hc.cy - R, 0), y1 = min(hc.cy + R, g.ny - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) scan_shell_row(sorted, cs, g, hc, R, y, z, p, top);
    done = top.at(mean_k) <= covered_d2(g, hc, R);  // +inf bound: the whole grid has been inspected
  }
  if (done) md[__float_as_int(p.w)] = mean_distance<K>(top, mean_k);
  else left_list[atomicAdd(left_count, 1u)] = (uint32_t)j;
}

// The `count` smallest values held by the 32 private lists of a warp, extracted in ascending order by repeated
// warp-min.  Returns the last one; s = sum of the f32 square roots of all but the first, accumulated in f64 in
// ascending order (PCL's dist_sum); lane 0 also stores the values to out[] when it is not null.
template <int K>
__device__ __forceinline__ float warp_extract(TopK<K> b, int count, int lane, double& s, float* out) {
  const unsigned FULL = 0xffffffffu;
  float kth = CUDART_INF_F;
  s = 0.0;
  for (int t = 0; t < count; ++t) {
    const unsigned hb = __float_as_uint(b.a[0]);  // non-negative floats order like their bit patterns
    const unsigned mn = __reduce_min_sync(FULL, hb);
    const int win = __ffs(__ballot_sync(FULL, hb == mn)) - 1;
    if (lane == win) {
#pragma unroll
      for (int j = 0; j + 1 < K; ++j) b.a[j] = b.a[j + 1];
      b.a[K - 1] = CUDART_INF_F;
    }
    kth = __uint_as_float(mn);
    if (t >= 1) s += (double)__fsqrt_rn(kth);
    if (out != nullptr && lane == 0) out[t] = kth;
  }
  return kth;
}

// warp per point: the lanes share the rows of every shell.  A point still uncertified after SOR_WARP_SHELLS shells
// is isolated; it goes to the exhaustive kernel below.
template <int K>
__global__ void __launch_bounds__(256)
sor_left_kernel(const float4* __restrict__ sorted, const uint32_t* __restrict__ cs, const GridParams* __restrict__ gp,
                int mean_k, float* __restrict__ md, const uint32_t* __restrict__ left_list,
                const uint32_t* __restrict__ left_count, uint32_t* __restrict__ brute_list,
                uint32_t* __restrict__ brute_count) {
  __shared__ GridParams g;
  if (threadIdx.x == 0) g = *gp;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned n_left = *left_count;
  for (unsigned w = warp; w < n_left; w += n_warps) {
    const uint32_t j = left_list[w];
    const float4 p = __ldg(sorted + j);
    const HomeCell hc = home_cell(g, p);
    TopK<K> top;
    top.init();
    double s = 0.0;
    bool done = false;
    for (int R = 0; R < SOR_WARP_SHELLS && !done; ++R) {
      const int side = 2 * R + 1, rows = side * side;
      for (int r = lane; r < rows; r += 32) {
        const int z = hc.cz + r / side - R, y = hc.cy + r % side - R;
        if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
        scan_shell_row(sorted, cs, g, hc, R, y, z, p, top);
      }
      const float kth = warp_extract<K>(top, mean_k + 1, lane, s, nullptr);
      done = kth <= covered_d2(g, hc, R);  // +inf bound: the whole grid has been inspected
    }
    if (lane == 0) {
      if (done) md[__float_as_int(p.w)] = (float)(s / (double)mean_k);
      else brute_list[atomicAdd(brute_count, 1u)] = j;
    }
  }
}

// block per isolated point: every point of the cloud is inspected once (coalesced), the per-thread lists are merged
// per warp and then across the warps.
constexpr int BR_BATCH = 8;  // independent loads in flight per thread
template <int K, int BR_THREADS>
__global__ void __launch_bounds__(BR_THREADS)
sor_brute_kernel(const float4* __restrict__ sorted, const GridParams* __restrict__ gp, int mean_k, float* __restrict__ md,
                 const uint32_t* __restrict__ brute_list, const uint32_t* __restrict__ brute_count) {
  __shared__ float sh[BR_THREADS / 32][32];
  const int n_points = gp->n_points;
  const unsigned n_brute = *brute_count;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (unsigned e = blockIdx.x; e < n_brute; e += gridDim.x) {
    const float4 p = __ldg(sorted + brute_list[e]);
    TopK<K> top;
    top.init();
    for (int base = 0; base < n_points; base += BR_THREADS * BR_BATCH) {
      float4 q[BR_BATCH];
#pragma unroll
      for (int u = 0; u < BR_BATCH; ++u) {
        const int t = base + u * BR_THREADS + (int)threadIdx.x;
        q[u] = t < n_points ? __ldg(sorted + t) : make_float4(CUDART_INF_F, 0.f, 0.f, 0.f);  // d2 = +inf: never kept
      }
#pragma unroll
      for (int u = 0; u < BR_BATCH; ++u) top.push(sor_d2(p, q[u]));
    }
    double s;
    for (int q = lane; q < 32; q += 32) sh[w][q] = CUDART_INF_F;
    __syncwarp();
    warp_extract<K>(top, mean_k + 1, lane, s, sh[w]);
    __syncthreads();
    if (w == 0) {
      TopK<K> m;
      m.init();
      if (lane < BR_THREADS / 32) {
#pragma unroll
        for (int q = 0; q < K; ++q) m.a[q] = sh[lane][q];  // ascending; entries past mean_k are +inf
      }
      warp_extract<K>(m, mean_k + 1, lane, s, nullptr);
      if (lane == 0) md[__float_as_int(p.w)] = (float)(s / (double)mean_k);
    }
    __syncthreads();
  }
}

// ---- stage 3: statistics of the mean distances and the keep flags (second pass of the PCL filter) --------
struct SorStats {
  double sum, sq_sum, mean, stddev, threshold;
  int borderline;
  unsigned n_left, n_brute;
};

constexpr int ST_THREADS = 256;
__global__ void __launch_bounds__(ST_THREADS)
sor_partial_kernel(const float* __restrict__ md, int n, double* __restrict__ partial) {
  __shared__ double sh[2][ST_THREADS / 32];
  double s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float d = md[i];
    s += (double)d;
    q += (double)(d * d);  // PCL squares in f32 before widening
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = s; sh[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < ST_THREADS / 32; ++k) { ts += sh[0][k]; tq += sh[1][k]; }
    partial[2 * blockIdx.x] = ts;
    partial[2 * blockIdx.x + 1] = tq;
  }
}
__global__ void sor_threshold_kernel(const double* __restrict__ partial, int n_partial, int n, double std_mul,
                                     const uint32_t* __restrict__ left_count, SorStats* __restrict__ st) {
  // one warp, fixed order: lane l adds partials l, l+32, ... then a shuffle tree
  const int lane = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int k = lane; k < n_partial; k += 32) { s += partial[2 * k]; q += partial[2 * k + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  const double valid = (double)n;
  const double mean = s / valid;
  const double variance = (q - s * s / valid) / (valid - 1.0);
  const double stddev = sqrt(variance);
  st->sum = s; st->sq_sum = q; st->mean = mean; st->stddev = stddev;
  st->threshold = mean + std_mul * stddev;
  st->borderline = 0;
  st->n_left = left_count[0];
  st->n_brute = left_count[1];
}
__global__ void __launch_bounds__(256)
sor_flag_kernel(const float* __restrict__ md, int n, SorStats* __restrict__ st, uint32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double thr = st->threshold;
  const double d = (double)md[i];
  flag[i] = (d > thr) ? 0u : 1u;  // removed iff distances[i] > distance_threshold (negative_ = false)
  if (fabs(d - thr) <= 1e-9 * fabs(thr)) atomicAdd(&st->borderline, 1);
}

template <int K>
cudaError_t launch_sor_knn(Ctx* c, const float4* sorted, const uint32_t* cs, const GridParams* d_gp, int n_points,
                           int rmax, int mean_k, float* md, uint32_t* left_list, uint32_t* left_count) {
  uint32_t* brute_list = left_list + n_points;
  uint32_t* brute_count = left_count + 1;
  sor_knn_kernel<K><<<div_up(n_points, SOR_THREADS), SOR_THREADS, 0, c->stream>>>(sorted, cs, d_gp, rmax, mean_k, md,
                                                                                  left_list, left_count);
  sor_left_kernel<K><<<c->sm_count * 2, 256, 0, c->stream>>>(sorted, cs, d_gp, mean_k, md, left_list, left_count,
                                                             brute_list, brute_count);
  constexpr int BR_THREADS = K <= 16 ? 1024 : 512;  // register budget of the K-entry private lists
  sor_brute_kernel<K, BR_THREADS><<<c->sm_count, BR_THREADS, 0, c->stream>>>(sorted, d_gp, mean_k, md, brute_list, brute_count);
  c->launches += 3;
  return cudaGetLastError();
}

}  // namespace

// Everything after the argument checks of liogpu_publish_local_map.  h_yaw: the 3x4 yaw-frame matrix and the crop
// limits, computed by the caller on the host (12 + 4 floats).  The result is left in *result (device, packed).
int publish_local_map_dev(Ctx* c, const float4* const* d_srcs, const int* d_offs, int k, const float* d_poses6,
                          float* d_T12, long long total, const float* h_yaw16, const liogpu_local_map_params* prm,
                          const float4** result, int* n_result, liogpu_local_map_info* info) {
  YawFrame yf;
  for (int q = 0; q < 12; ++q) yf.m[q] = h_yaw16[q];
  yf.xmin = h_yaw16[12]; yf.xmax = h_yaw16[13]; yf.ymin = h_yaw16[14]; yf.ymax = h_yaw16[15];
  const int n = (int)total;
  LIOGPU_CUDA_OK(c, c->map_raw4.reserve((size_t)n * sizeof(float4)));
  LIOGPU_CUDA_OK(c, c->lm_flag.reserve((size_t)n * sizeof(uint32_t)));
  LIOGPU_CUDA_OK(c, c->lm_pos.reserve((size_t)n * sizeof(uint32_t)));
  LIOGPU_CUDA_OK(c, c->lm_a.reserve((size_t)n * sizeof(float4)));
  LIOGPU_CUDA_OK(c, c->lm_stats.reserve(4096 + 2 * 1024 * sizeof(double)));
  uint32_t* d_total = c->lm_stats.as<uint32_t>();          // [0]: compaction total, [1]: leftover count
  SorStats* d_st = reinterpret_cast<SorStats*>((char*)c->lm_stats.p + 256);
  double* d_partial = reinterpret_cast<double*>((char*)c->lm_stats.p + 4096);
  uint32_t* h_total = reinterpret_cast<uint32_t*>((char*)c->h_pinned + 8192);
  SorStats* h_st = reinterpret_cast<SorStats*>((char*)c->h_pinned + 8448);

  // stage 1
  LIOGPU_CUDA_OK(c, launch_pose_table(c, d_poses6, k, d_T12));
  lmap_transform_crop_kernel<<<div_up(total, 256), 256, 0, c->stream>>>(d_srcs, d_offs, k, d_T12, yf, total,
                                                                        c->map_raw4.as<float4>(), c->lm_flag.as<uint32_t>());
  c->launches++;
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, exclusive_scan_u32(c, c->lm_flag.as<uint32_t>(), c->lm_pos.as<uint32_t>(), n, d_total));
  lmap_compact_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(c->map_raw4.as<float4>(), c->lm_flag.as<uint32_t>(),
                                                             c->lm_pos.as<uint32_t>(), n, c->lm_a.as<float4>());
  c->launches++;
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_total, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  int n_c = (int)h_total[0];
  const float4* cur = c->lm_a.as<float4>();
  info->n_concat = n;
  info->n_cropped = n_c;
  info->n_after_sor = n_c;

  // stage 2 + 3.  With fewer than meanK+1 points PCL's search returns short, every distance is 0 and
  // nothing is removed (statistical_outlier_removal.hpp) — the filter is the identity.
  if (prm->use_removing_outliers && n_c >= prm->mean_k + 1) {
    const int mean_k = prm->mean_k;
    float cell = prm->sor_cell_size > 0.f ? prm->sor_cell_size : 0.4f;
    GridParams g;
    int rc = grid_build_core(c, cur, n_c, cell, 1.0f, 1.0f, c->sor_setup, c->sor_sorted, c->sor_cell_start, g);
    if (rc) return rc;
    LIOGPU_CUDA_OK(c, c->lm_md.reserve((size_t)n_c * sizeof(float)));
    LIOGPU_CUDA_OK(c, c->lm_left.reserve((size_t)n_c * 2 * sizeof(uint32_t)));
    LIOGPU_CUDA_OK(c, cudaMemsetAsync(d_total + 1, 0, 2 * sizeof(uint32_t), c->stream));
    const float4* sorted = c->sor_sorted.as<float4>();
    const uint32_t* cs = c->sor_cell_start.as<uint32_t>();
    const GridParams* d_gp = c->sor_setup.as<GridParams>();
    float* md = c->lm_md.as<float>();
    uint32_t* left = c->lm_left.as<uint32_t>();
    const int rmax = 2;
    cudaError_t e;
    if (mean_k + 1 <= 6) e = launch_sor_knn<6>(c, sorted, cs, d_gp, g.n_points, rmax, mean_k, md, left, d_total + 1);
    else if (mean_k + 1 <= 11) e = launch_sor_knn<11>(c, sorted, cs, d_gp, g.n_points, rmax, mean_k, md, left, d_total + 1);
    else if (mean_k + 1 <= 16) e = launch_sor_knn<16>(c, sorted, cs, d_gp, g.n_points, rmax, mean_k, md, left, d_total + 1);
    else e = launch_sor_knn<32>(c, sorted, cs, d_gp, g.n_points, rmax, mean_k, md, left, d_total + 1);
    LIOGPU_CUDA_OK(c, e);
    const int n_partial = 1024 < div_up(n_c, ST_THREADS) ? 1024 : div_up(n_c, ST_THREADS);
    sor_partial_kernel<<<n_partial, ST_THREADS, 0, c->stream>>>(md, n_c, d_partial);
    sor_threshold_kernel<<<1, 32, 0, c->stream>>>(d_partial, n_partial, n_c, (double)prm->stddev_threshold, d_total + 1, d_st);
    sor_flag_kernel<<<div_up(n_c, 256), 256, 0, c->stream>>>(md, n_c, d_st, c->lm_flag.as<uint32_t>());
    c->launches += 3;
    LIOGPU_CUDA_OK(c, cudaGetLastError());
    LIOGPU_CUDA_OK(c, exclusive_scan_u32(c, c->lm_flag.as<uint32_t>(), c->lm_pos.as<uint32_t>(), n_c, d_total));
    LIOGPU_CUDA_OK(c, c->lm_b.reserve((size_t)n_c * sizeof(float4)));
    lmap_compact_kernel<<<div_up(n_c, 256), 256, 0, c->stream>>>(cur, c->lm_flag.as<uint32_t>(), c->lm_pos.as<uint32_t>(),
                                                                 n_c, c->lm_b.as<float4>());
    c->launches++;
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_total, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_st, d_st, sizeof(SorStats), cudaMemcpyDeviceToHost, c->stream));
    LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    n_c = (int)h_total[0];
    cur = c->lm_b.as<float4>();
    info->n_after_sor = n_c;
    info->sor_mean = h_st->mean;
    info->sor_stddev = h_st->stddev;
    info->sor_threshold = h_st->threshold;
    info->sor_borderline = h_st->borderline;
    info->sor_leftover = (int)h_st->n_left;
    info->sor_exhaustive = (int)h_st->n_brute;
  }

  // stage 4
  if (prm->use_down_sampling && n_c > 0) {
    int m = 0;
    bool overflow = false;
    int rc = voxel_downsample_dev(c, cur, n_c, prm->local_mapping_surf_leaf_size, c->lm_out, &m, &overflow);
    if (rc) return rc;
    info->leaf_overflow = overflow ? 1 : 0;
    n_c = m;
    cur = c->lm_out.as<float4>();
  }
  info->n_out = n_c;
  *result = cur;
  *n_result = n_c;
  return LIOGPU_OK;
}

}  // namespace liogpu
