// icp.cu — the loop-closure registration: pcl::IterativeClosestPoint<PointXYZI, PointXYZI>::align as
// configured at mapOptmization.cpp:1111-1121 (also :1203-1213) and icp.getFitnessScore() (:1123); SURVEY §8 row f3.
//
// Per ICP iteration (PCL registration/impl/icp.hpp), all on device, no host round trip inside a chunk:
//   icp_nn_kernel      thread per source point: move it by the previous iteration's transformation_, exact
//                      nearest target point on the sorted grid, shells 0..2 (ties: lower target index)
//   icp_left_kernel    warp per point that needs shells 3..5
//   icp_brute_kernel   block per isolated point: one pass over the target
//   icp_reduce_kernel  f64 sums over the correspondences (d^2 <= max^2): source / target centroids, cross
//                      products, squared distances; the last block runs Eigen::umeyama (3x3 SVD by one-sided
//                      Jacobi in f64), final_transformation_ = transformation_ * final_transformation_ and
//                      DefaultConvergenceCriteria::hasConverged.
// The search stops growing once every uninspected cell is farther than the correspondence distance (such a point
// has no correspondence).  getFitnessScore: the same search with no distance limit on the source moved by the final
// transformation, f64 mean of the squared distances.
#include "common.cuh"

#include <stdlib.h>
#include "shell_search.cuh"

namespace liogpu {

namespace {

constexpr int ICP_SUMS = 17;  // 3 src + 3 tgt + 9 cross + mse + count

struct IcpDevState {
  float T_inc[16];   // transformation_ of the last finished iteration (row-major 4x4)
  float finalT[16];  // final_transformation_
  double prev_mse, last_mse, fitness;
  double max_d2, rot_thr, trans_thr, rel_thr, abs_thr;
  int iter, done, state, n_corr, max_iter, fit_nr;
  unsigned n_left, n_brute, ticket, pad;
};

struct IcpArgs {
  const float4* src;        // original source
  float4* cur;              // input_transformed
  int ns;
  const float4* tgt;        // target, original order
  const float4* sorted;     // target in grid order, w = bits(original index)
  const uint32_t* cs;
  const GridParams* gp;
  IcpDevState* st;
  int* nn_idx;
  float* nn_d2;
  uint32_t* left_list;      // [ns] shells 3.., then [ns] exhaustive
  double* partials;         // [reduce blocks][ICP_SUMS]
  int fitness;              // 1: getFitnessScore pass (source moved by finalT, no distance limit)
  int all_warp;             // 1: every source point goes straight to the warp-per-point search (small sources: a
                            //    thread-per-point pass would leave most of the GPU idle and is latency bound)
};

// nearest point: (bits(d^2) << 32 | target index), so that min() also breaks ties toward the lower index
struct Best1 {
  unsigned long long key;
  __device__ __forceinline__ void init() { key = (0x7f800000ULL << 32) | 0xffffffffULL; }  // +inf, index -1
  __device__ __forceinline__ void push(float d2, float w) {
    const unsigned long long k = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(w);
    key = k < key ? k : key;
  }
  __device__ __forceinline__ float d2() const { return __uint_as_float((unsigned)(key >> 32)); }
  __device__ __forceinline__ int idx() const { return (int)(unsigned)(key & 0xffffffffULL); }
};

// pcl::transformPointCloud with a Matrix4f (PCL >= 1.10 detail::Transformer<float>::se3)
__device__ __forceinline__ float4 se3(const float* T, const float4 p) {
  float4 q;
  q.x = p.x * T[0] + (p.y * T[1] + (p.z * T[2] + T[3]));
  q.y = p.x * T[4] + (p.y * T[5] + (p.z * T[6] + T[7]));
  q.z = p.x * T[8] + (p.y * T[9] + (p.z * T[10] + T[11]));
  q.w = p.w;
  return q;
}

// 0: keep growing, 1: nearest point certified, 2: nothing within the correspondence distance
__device__ __forceinline__ int nn_verdict(const Best1& b, float cov2, double max_d2) {
  if (b.d2() <= cov2) return 1;
  if ((double)cov2 > max_d2) return 2;
  return 0;
}

constexpr int NN_THREADS = 128;
constexpr int NN_THREAD_SHELLS = 3;  // shells 0..2 per thread
constexpr int NN_WARP_SHELLS = 6;    // shells 0..5 per warp

__global__ void __launch_bounds__(NN_THREADS)
icp_nn_kernel(const IcpArgs A) {
  __shared__ GridParams g;
  __shared__ float sT[12];
  __shared__ int s_skip, s_move;
  __shared__ double s_max_d2;
  if (threadIdx.x == 0) {
    g = *A.gp;
    s_skip = A.fitness ? 0 : A.st->done;
    s_move = A.fitness ? 1 : (A.st->iter > 0);
    s_max_d2 = A.fitness ? CUDART_INF : A.st->max_d2;
  }
  if (threadIdx.x < 12) sT[threadIdx.x] = A.fitness ? A.st->finalT[threadIdx.x] : A.st->T_inc[threadIdx.x];
  __syncthreads();
  if (s_skip) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.ns) return;
  float4 p = A.fitness ? A.src[i] : A.cur[i];
  if (s_move) {
    p = se3(sT, p);
    A.cur[i] = p;
  }
  if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) { A.nn_idx[i] = -1; A.nn_d2[i] = CUDART_INF_F; return; }
  const HomeCell hc = home_cell(g, p);
  Best1 best;
  best.init();
  int verdict = 0;
  for (int R = 0; R < NN_THREAD_SHELLS && verdict == 0; ++R) {
    const int z0 = max(hc.cz - R, 0), z1 = min(hc.cz + R, g.nz - 1);
    const int y0 = max(hc.cy - R, 0), y1 = min(hc.cy + R, g.ny - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) scan_shell_row(A.sorted, A.cs, g, hc, R, y, z, p, best);
    verdict = nn_verdict(best, covered_d2(g, hc, R), s_max_d2);
  }
  if (verdict == 0) {
    A.left_list[atomicAdd(&A.st->n_left, 1u)] = (uint32_t)i;
  } else {
    A.nn_idx[i] = verdict == 1 ? best.idx() : -1;
    A.nn_d2[i] = best.d2();
  }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u < v ? u : v;
  }
  return v;
}

__global__ void __launch_bounds__(256)
icp_left_kernel(const IcpArgs A) {
  __shared__ GridParams g;
  __shared__ float sT[12];
  if (threadIdx.x == 0) g = *A.gp;
  if (threadIdx.x < 12) sT[threadIdx.x] = A.fitness ? A.st->finalT[threadIdx.x] : A.st->T_inc[threadIdx.x];
  __syncthreads();
  if (!A.fitness && A.st->done) return;
  const double max_d2 = A.fitness ? CUDART_INF : A.st->max_d2;
  const bool move = A.all_warp && (A.fitness || A.st->iter > 0);
  const int lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned n_left = A.all_warp ? (unsigned)A.ns : A.st->n_left;
  for (unsigned w = warp; w < n_left; w += n_warps) {
    const uint32_t i = A.all_warp ? w : A.left_list[w];
    float4 p = (A.all_warp && A.fitness) ? A.src[i] : A.cur[i];
    if (move) {
      p = se3(sT, p);
      if (lane == 0) A.cur[i] = p;
    }
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {
      if (lane == 0) { A.nn_idx[i] = -1; A.nn_d2[i] = CUDART_INF_F; }
      continue;
    }
    const HomeCell hc = home_cell(g, p);
    Best1 best;
    best.init();
    int verdict = 0;
    Best1 all;
    all.init();
    for (int R = 0; R < NN_WARP_SHELLS && verdict == 0; ++R) {
      const int side = 2 * R + 1, rows = side * side;
      for (int r = lane; r < rows; r += 32) {
        const int z = hc.cz + r / side - R, y = hc.cy + r % side - R;
        if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
        scan_shell_row(A.sorted, A.cs, g, hc, R, y, z, p, best);
      }
      all.key = warp_min_u64(best.key);
      verdict = nn_verdict(all, covered_d2(g, hc, R), max_d2);
    }
    if (lane == 0) {
      if (verdict == 0) {
        A.left_list[A.ns + atomicAdd(&A.st->n_brute, 1u)] = i;
      } else {
        A.nn_idx[i] = verdict == 1 ? all.idx() : -1;
        A.nn_d2[i] = all.d2();
      }
    }
  }
}

constexpr int BRT = 512;
__global__ void __launch_bounds__(BRT)
icp_brute_kernel(const IcpArgs A) {
  __shared__ unsigned long long sh[BRT / 32];
  if (!A.fitness && A.st->done) return;
  const double max_d2 = A.fitness ? CUDART_INF : A.st->max_d2;
  const int n_points = A.gp->n_points;
  const unsigned n_brute = A.st->n_brute;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (unsigned e = blockIdx.x; e < n_brute; e += gridDim.x) {
    const uint32_t i = A.left_list[A.ns + e];
    const float4 p = A.cur[i];
    Best1 best;
    best.init();
    for (int base = 0; base < n_points; base += BRT * 8) {
      float4 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = base + u * BRT + (int)threadIdx.x;
        q[u] = t < n_points ? __ldg(A.sorted + t) : make_float4(CUDART_INF_F, 0.f, 0.f, __int_as_float(-1));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) best.push(sor_d2(p, q[u]), q[u].w);
    }
    const unsigned long long wm = warp_min_u64(best.key);
    if (lane == 0) sh[w] = wm;
    __syncthreads();
    if (w == 0) {
      unsigned long long v = lane < BRT / 32 ? sh[lane] : ~0ULL;
      v = warp_min_u64(v);
      if (lane == 0) {
        Best1 b;
        b.key = v;
        const bool ok = (double)b.d2() <= max_d2;  // exhaustive: always the true nearest point
        A.nn_idx[i] = ok ? b.idx() : -1;
        A.nn_d2[i] = b.d2();
      }
    }
    __syncthreads();
  }
}

// ---- Eigen::umeyama (no scaling) from the f64 sums; same operation order as the oracle's restatement ----
__device__ void umeyama_rotation_dev(const float sigma[9], float R[9]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = (double)sigma[3 * i + j];
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
        if (gamma == 0.0 || fabs(gamma) <= 1e-17 * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int i = 0; i < 3; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - sn * aq; A[i][q] = sn * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sv[3], U[3][3];
  for (int j = 0; j < 3; ++j) sv[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  int lo = 0;
  if (sv[1] < sv[lo]) lo = 1;
  if (sv[2] < sv[lo]) lo = 2;
  const int a = (lo + 1) % 3, b = (lo + 2) % 3;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) U[i][j] = sv[j] > 0.0 ? A[i][j] / sv[j] : 0.0;
  if (!(sv[lo] > 1e-12 * (sv[a] > sv[b] ? sv[a] : sv[b]))) {  // rank-deficient: complete U right-handed
    U[0][lo] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
    U[1][lo] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
    U[2][lo] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
  }
  const double detU = U[0][0] * (U[1][1] * U[2][2] - U[1][2] * U[2][1]) - U[0][1] * (U[1][0] * U[2][2] - U[1][2] * U[2][0]) +
                      U[0][2] * (U[1][0] * U[2][1] - U[1][1] * U[2][0]);
  const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                      V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
  const double d = detU * detV < 0.0 ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double r = U[i][a] * V[j][a] + U[i][b] * V[j][b] + d * (U[i][lo] * V[j][lo]);
      R[3 * i + j] = (float)r;
    }
}

__device__ void icp_finalize(IcpDevState* st, const double* S) {
  const int cnt = (int)S[16];
  st->n_corr = cnt;
  if (cnt < 3) {  // min_number_correspondences_
    st->state = 5;
    st->done = 1;
    return;
  }
  const double inv_n = 1.0 / (double)cnt;
  const double mse = S[15] / (double)cnt;
  st->last_mse = mse;
  float src_mean[3], dst_mean[3], sigma[9], R[9], T[16];
  for (int a = 0; a < 3; ++a) { src_mean[a] = (float)(S[a] * inv_n); dst_mean[a] = (float)(S[3 + a] * inv_n); }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) sigma[3 * i + j] = (float)((S[6 + 3 * i + j] - S[3 + i] * S[j] * inv_n) * inv_n);
  umeyama_rotation_dev(sigma, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
    T[4 * i + 3] = dst_mean[i] - (R[3 * i] * src_mean[0] + R[3 * i + 1] * src_mean[1] + R[3 * i + 2] * src_mean[2]);
  }
  T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
  float F[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      F[4 * i + j] = T[4 * i] * st->finalT[j] + T[4 * i + 1] * st->finalT[4 + j] + T[4 * i + 2] * st->finalT[8 + j] +
                     T[4 * i + 3] * st->finalT[12 + j];
  for (int q = 0; q < 16; ++q) { st->finalT[q] = F[q]; st->T_inc[q] = T[q]; }
  const int it = ++st->iter;
  // DefaultConvergenceCriteria::hasConverged
  int state = 0;
  if (it >= st->max_iter) {
    state = 1;
  } else {
    const float tr = T[0] + T[5] + T[10] - 1;
    const double cos_angle = 0.5 * tr;
    const float tsq = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
    const double translation_sqr = tsq;
    if (cos_angle >= st->rot_thr && translation_sqr <= st->trans_thr) state = 2;
    else if (fabs(mse - st->prev_mse) < st->abs_thr) state = 3;
    else if (fabs(mse - st->prev_mse) / st->prev_mse < st->rel_thr) state = 4;
    else st->prev_mse = mse;
  }
  st->state = state;
  if (state != 0) st->done = 1;
}

constexpr int RD_THREADS = 256;
__global__ void __launch_bounds__(RD_THREADS)
icp_reduce_kernel(const IcpArgs A) {
  __shared__ double sh[RD_THREADS / 32][ICP_SUMS];
  __shared__ bool s_last;
  IcpDevState* st = A.st;
  if (!A.fitness && st->done) return;
  const double max_d2 = A.fitness ? CUDART_INF : st->max_d2;
  double acc[ICP_SUMS];
#pragma unroll
  for (int q = 0; q < ICP_SUMS; ++q) acc[q] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.ns; i += gridDim.x * blockDim.x) {
    const int j = A.nn_idx[i];
    const float d2 = A.nn_d2[i];
    if (j < 0 || (double)d2 > max_d2) continue;
    acc[15] += (double)d2;
    acc[16] += 1.0;
    if (!A.fitness) {
      const float4 s = A.cur[i], t = A.tgt[j];
      const double sv[3] = {s.x, s.y, s.z}, tv[3] = {t.x, t.y, t.z};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        acc[a] += sv[a];
        acc[3 + a] += tv[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[6 + 3 * a + b] += tv[a] * sv[b];
      }
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < ICP_SUMS; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < ICP_SUMS) {
    double v = 0.0;
    for (int k = 0; k < RD_THREADS / 32; ++k) v += sh[k][threadIdx.x];
    A.partials[(size_t)blockIdx.x * ICP_SUMS + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {  // fixed order: eight interleaved slices of the blocks (their loads in flight together), then the slices
    const int q = threadIdx.x & 31, slice = threadIdx.x >> 5;
    double v = 0.0;
    if (q < ICP_SUMS) {
      for (unsigned b = slice; b < gridDim.x; b += RD_THREADS / 32) v += A.partials[(size_t)b * ICP_SUMS + q];
      sh[slice][q] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < ICP_SUMS) {
    double v = 0.0;
    for (int k = 0; k < RD_THREADS / 32; ++k) v += sh[k][threadIdx.x];
    sh[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (A.fitness) {
      st->fit_nr = (int)sh[0][16];
      st->fitness = sh[0][16] > 0.0 ? sh[0][15] / sh[0][16] : DBL_MAX;
    } else {
      icp_finalize(st, sh[0]);
    }
    st->ticket = 0;
    st->n_left = 0;
    st->n_brute = 0;
  }
}

}  // namespace

// source / target: packed float4 on device.  The target's grid index is built into the outlier filter's buffers
// (publishLocalMap and the loop closure never run concurrently on one context).
int icp_align_dev(Ctx* c, const float4* src, int ns, const float4* tgt, int nt, const liogpu_icp_params* prm,
                  float final_T[16], liogpu_icp_info* info) {
  GridParams g;
  const float cell = prm->cell_size > 0.f ? prm->cell_size : 0.5f;  // measured: 0.5 m 1.33 ms, 1 m 1.58 ms, 1.5 m 1.96 ms
  int rc = grid_build_core(c, tgt, nt, cell, 1.0f, 1.0f, c->sor_setup, c->sor_sorted, c->sor_cell_start, g);
  if (rc) return rc;
  const int rblocks = div_up(ns, RD_THREADS) < 256 ? div_up(ns, RD_THREADS) : 256;
  LIOGPU_CUDA_OK(c, c->lm_a.reserve((size_t)ns * sizeof(float4)));
  LIOGPU_CUDA_OK(c, c->lm_flag.reserve((size_t)ns * sizeof(int)));
  LIOGPU_CUDA_OK(c, c->lm_md.reserve((size_t)ns * sizeof(float)));
  LIOGPU_CUDA_OK(c, c->lm_left.reserve((size_t)ns * 2 * sizeof(uint32_t)));
  LIOGPU_CUDA_OK(c, c->lm_stats.reserve(4096 + (size_t)256 * ICP_SUMS * sizeof(double)));
  IcpArgs A;
  A.src = src; A.cur = c->lm_a.as<float4>(); A.ns = ns;
  A.tgt = tgt; A.sorted = c->sor_sorted.as<float4>(); A.cs = c->sor_cell_start.as<uint32_t>();
  A.gp = c->sor_setup.as<GridParams>();
  A.st = reinterpret_cast<IcpDevState*>((char*)c->lm_stats.p + 1024);
  A.nn_idx = c->lm_flag.as<int>(); A.nn_d2 = c->lm_md.as<float>();
  A.left_list = c->lm_left.as<uint32_t>();
  A.partials = reinterpret_cast<double*>((char*)c->lm_stats.p + 4096);
  A.fitness = 0;
  A.all_warp = (ns <= 65536 && !getenv("LIOGPU_ICP_THREAD_PASS")) ? 1 : 0;
  IcpDevState* h = reinterpret_cast<IcpDevState*>((char*)c->h_pinned + 12288);
  memset(h, 0, sizeof(IcpDevState));
  for (int q = 0; q < 16; ++q) h->T_inc[q] = h->finalT[q] = (q % 5 == 0) ? 1.f : 0.f;
  h->prev_mse = DBL_MAX;
  h->max_d2 = (double)prm->max_correspondence_distance * (double)prm->max_correspondence_distance;
  h->rot_thr = 1.0 - prm->transformation_epsilon;   // transformation_rotation_epsilon_ unset (icp.hpp)
  h->trans_thr = prm->transformation_epsilon;
  h->rel_thr = prm->euclidean_fitness_epsilon;
  h->abs_thr = 1e-12;
  h->max_iter = prm->max_iterations;
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(A.st, h, sizeof(IcpDevState), cudaMemcpyHostToDevice, c->stream));
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(A.cur, src, (size_t)ns * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
  const int CHUNK = 10;
  int launched = 0;
  auto enqueue_iteration = [&](const IcpArgs& a) {
    if (!a.all_warp) icp_nn_kernel<<<div_up(ns, NN_THREADS), NN_THREADS, 0, c->stream>>>(a);
    icp_left_kernel<<<a.all_warp ? c->sm_count * 8 : c->sm_count * 2, 256, 0, c->stream>>>(a);
    icp_brute_kernel<<<c->sm_count, BRT, 0, c->stream>>>(a);
    icp_reduce_kernel<<<rblocks, RD_THREADS, 0, c->stream>>>(a);
    c->launches += a.all_warp ? 3 : 4;
  };
  for (;;) {
    const int todo = (prm->max_iterations - launched) < CHUNK ? (prm->max_iterations - launched) : CHUNK;
    for (int it = 0; it < todo; ++it) enqueue_iteration(A);
    launched += todo;
    LIOGPU_CUDA_OK(c, cudaGetLastError());
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h, A.st, sizeof(IcpDevState), cudaMemcpyDeviceToHost, c->stream));
    LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (h->done || launched >= prm->max_iterations) break;
  }
  // getFitnessScore
  IcpArgs F = A;
  F.fitness = 1;
  enqueue_iteration(F);
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h, A.st, sizeof(IcpDevState), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
  for (int q = 0; q < 16; ++q) final_T[q] = h->finalT[q];
  info->iterations = h->iter;
  info->convergence_state = h->state;
  info->converged = (h->state >= 1 && h->state <= 4) ? 1 : 0;
  info->n_correspondences = h->n_corr;
  info->fitness_score = h->fitness;
  info->last_mse = h->last_mse;
  info->gpu_ms = c->last_ms;
  return LIOGPU_OK;
}

}  // namespace liogpu
