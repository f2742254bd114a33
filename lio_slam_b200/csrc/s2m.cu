// s2m.cu — the scan-to-map hot loop: surfOptimization + combineOptimizationCoeffs + LMOptimization
// (mapOptmization.cpp:1618-1687, 1689-1700, 1702-1837) as ONE kernel per Gauss-Newton iteration.
//
// s2m_iter_kernel, one thread per scan point:
//   1. pointAssociateToMap (:841-847) with the 3x4 transform built from the device-resident pose
//   2. exact 5-NN inside the 1 m gate on the sorted grid (grid.cu): rows of x-adjacent cells are
//      contiguous ranges of map_sorted; rows are visited centre-out and pruned against the current
//      5th-best distance; distances are FLANN L2_Simple in f32 without FMA; ties -> lower map index
//   3. 5x3 column-pivoted Householder plane fit, validity, weight s, coefficient (:1633-1684)
//   4. Jacobian row (:1760-1778) staged in shared memory; the block reduces the 27 sums of
//      A^T A (upper triangle) and A^T b in FP64 (cv::gemm accumulates f32 products in double)
//   5. the last block to finish (threadfence + atomic ticket) adds the per-block partials in a fixed
//      order and runs the 6x6 tail of LMOptimization on device: QR solve, Jacobi eigen + matP on
//      iteration 0, degeneracy projection, pose update, convergence test (:1784-1835).
// The pose, matP, isDegenerate and the iteration counter stay in HBM (LmDevState); the host enqueues
// max_iter launches back to back and never reads anything until the loop is over — a launch that
// finds `done` set returns immediately.  Compaction (:1689-1700) is unnecessary: rejected points
// contribute exact zeros and the FP64 sums do not depend on the order.
//
// Algorithmic HBM bytes per launch: 16 B query + 5 x 16 B neighbours = 96 B per scan point.
#include "common.cuh"
#include "pose_math.cuh"

namespace liogpu {

constexpr int S2M_THREADS = 256;
constexpr int S2M_SUMS = 32;  // 21 (upper AtA) + 6 (AtB) + nsel + ties, padded to 32

struct SurfDebugOut {
  int* nn_idx;          // n*5
  float* nn_d2;         // n*5
  float4* coeff;        // n
  unsigned char* flag;  // n
  unsigned char* tie;   // n
};

#define LIOGPU_LT(da, ia, db, ib) ((da) < (db) || ((da) == (db) && (ia) < (ib)))

struct Top5 {
  float d0, d1, d2, d3, d4;
  int i0, i1, i2, i3, i4;
  float rej;  // best distance among candidates that are not in the top 5 (tie logging)
  __device__ __forceinline__ void init(float gate) {
    d0 = d1 = d2 = d3 = d4 = gate;
    i0 = i1 = i2 = i3 = i4 = -1;
    rej = FLT_MAX;
  }
  __device__ __forceinline__ void offer(float d, int id) {
    if (LIOGPU_LT(d, id, d4, i4)) {
      rej = fminf(rej, d4);
      d4 = d; i4 = id;
      if (LIOGPU_LT(d4, i4, d3, i3)) {
        swapf(d3, d4); swapi(i3, i4);
        if (LIOGPU_LT(d3, i3, d2, i2)) {
          swapf(d2, d3); swapi(i2, i3);
          if (LIOGPU_LT(d2, i2, d1, i1)) {
            swapf(d1, d2); swapi(i1, i2);
            if (LIOGPU_LT(d1, i1, d0, i0)) { swapf(d0, d1); swapi(i0, i1); }
          }
        }
      }
    } else {
      rej = fminf(rej, d);
    }
  }
};

// Exact 5 nearest map points of q within sqrt(gate_d2), ascending (d2, map index).
__device__ __forceinline__ void grid_knn5(const float4 q, const GridParams& g, const float4* __restrict__ map_sorted,
                                          const uint32_t* __restrict__ cell_start, Top5& t) {
  t.init(g.gate_d2);
  const float s2 = 2.0f * g.slack;
  const float reach = sqrtf(g.gate_d2) + s2;
  const int R = (int)ceilf(reach * g.inv_h);
  const int cz = (int)floorf((q.z - g.oz) * g.inv_h);
  const int cy = (int)floorf((q.y - g.oy) * g.inv_h);
  for (int kz = 0; kz <= 2 * R; ++kz) {
    const int dz = (kz & 1) ? -((kz + 1) >> 1) : (kz >> 1);  // 0,-1,+1,-2,+2,...
    const int z = cz + dz;
    if (z < 0 || z >= g.nz) continue;
    const float zlo = g.oz + (float)z * g.h;
    float gz = fmaxf(zlo - q.z, q.z - (zlo + g.h)) - s2;
    gz = fmaxf(gz, 0.f);
    const float gz2 = gz * gz;
    if (gz2 > t.d4) continue;
    for (int ky = 0; ky <= 2 * R; ++ky) {
      const int dy = (ky & 1) ? -((ky + 1) >> 1) : (ky >> 1);
      const int y = cy + dy;
      if (y < 0 || y >= g.ny) continue;
      const float ylo = g.oy + (float)y * g.h;
      float gy = fmaxf(ylo - q.y, q.y - (ylo + g.h)) - s2;
      gy = fmaxf(gy, 0.f);
      const float m2 = (gz2 + gy * gy) * 0.999999f;
      if (m2 > t.d4) continue;
      const float r = sqrtf(t.d4 - m2) * 1.000001f + s2;
      int xlo = (int)floorf((q.x - r - g.ox) * g.inv_h);
      int xhi = (int)floorf((q.x + r - g.ox) * g.inv_h);
      xlo = xlo < 0 ? 0 : xlo;
      xhi = xhi >= g.nx ? g.nx - 1 : xhi;
      if (xlo > xhi) continue;
      const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
      const uint32_t s = __ldg(cell_start + row + xlo);
      const uint32_t e = __ldg(cell_start + row + xhi + 1);
      for (uint32_t j = s; j < e; ++j) {
        const float4 p = __ldg(map_sorted + j);
        float d = q.x - p.x;
        float acc = d * d;            // FLANN L2_Simple: result = 0; result += diff*diff (x, y, z)
        d = q.y - p.y; acc = acc + d * d;
        d = q.z - p.z; acc = acc + d * d;
        t.offer(acc, __float_as_int(p.w));
      }
    }
  }
}

// ---- the 6x6 tail of LMOptimization, one thread (mapOptmization.cpp:1721-1835) ----
__device__ __noinline__ void lm_finalize(LmDevState* st, const double* sums) {
  const int it = st->iter;
  const int nsel = (int)sums[27];
  st->n_sel = nsel;
  st->tie_queries = (int)sums[28];
  st->nsel_hist[it] = nsel;
  // expand the upper triangle; AtA(a,b) and AtA(b,a) are the same f64 sum of the same products
  int p = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b) {
      st->JtJ[a * 6 + b] = sums[p];
      st->JtJ[b * 6 + a] = sums[p];
      ++p;
    }
  for (int a = 0; a < 6; ++a) st->Jtr[a] = sums[21 + a];
  bool conv = false;
  if (nsel >= 50) {  // :1721-1724 — below 50 the pose is untouched and the loop just repeats
    float AtA[36], AtB[6], X[6];
    for (int i = 0; i < 36; ++i) AtA[i] = (float)st->JtJ[i];
    for (int i = 0; i < 6; ++i) AtB[i] = (float)st->Jtr[i];
    solve6_qr(AtA, AtB, X);
    if (it == 0) {  // :1786-1808
      float E[6], V[36], V2[36], Vi[36];
      eigen6_jacobi(AtA, E, V);
      for (int i = 0; i < 36; ++i) V2[i] = V[i];
      int deg = 0;
      for (int i = 5; i >= 0; --i) {
        if (E[i] < 100.f) {
          for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.f;
          deg = 1;
        } else {
          break;
        }
      }
      st->degenerate = deg;
      inv6_lu(V, Vi);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += (double)Vi[i * 6 + k] * (double)V2[k * 6 + j];
          st->matP[i * 6 + j] = (float)s;
        }
    }
    if (st->degenerate) {  // :1810-1815
      float X2[6];
      for (int i = 0; i < 6; ++i) X2[i] = X[i];
      for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += (double)st->matP[i * 6 + k] * (double)X2[k];
        X[i] = (float)s;
      }
    }
    for (int i = 0; i < 6; ++i) st->pose[i] += X[i];
    const float r2d = 57.29578f;  // pcl::rad2deg(float)
    const float rx = X[0] * r2d, ry = X[1] * r2d, rz = X[2] * r2d;
    const float dr = (float)sqrt((double)rx * rx + (double)ry * ry + (double)rz * rz);
    const float tx = X[3] * 100, ty = X[4] * 100, tz = X[5] * 100;
    const float dt = (float)sqrt((double)tx * tx + (double)ty * ty + (double)tz * tz);
    st->delta_r = dr;
    st->delta_t = dt;
    conv = ((double)dr < 0.05) && ((double)dt < 0.05);
  }
  for (int i = 0; i < 6; ++i) st->pose_hist[it][i] = st->pose[i];
  st->iter = it + 1;
  if (nsel < 50) {
    // Nothing changed, so every remaining iteration of the reference's loop would redo identical work
    // and bail out at :1722 again (quirk q2): record them and stop instead of burning launches.
    for (int k = it + 1; k < st->max_iter; ++k) {
      for (int i = 0; i < 6; ++i) st->pose_hist[k][i] = st->pose[i];
      st->nsel_hist[k] = nsel;
    }
    st->iter = st->max_iter;
  }
  if (conv) { st->converged = 1; st->done = 1; }
  if (st->iter >= st->max_iter) st->done = 1;
}

// mode 0: LM iteration on the device state.  mode 1: one surfOptimization pass with per-point outputs
// (no state update); T_override (12 floats) replaces the pose-derived transform when non-null.
__global__ void __launch_bounds__(S2M_THREADS)
s2m_iter_kernel(const float4* __restrict__ scan, int nq, const float4* __restrict__ map4,
                const float4* __restrict__ map_sorted, const uint32_t* __restrict__ cell_start, const GridParams g,
                LmDevState* __restrict__ st, double* __restrict__ partials, unsigned* __restrict__ ticket,
                const float* __restrict__ T_override, SurfDebugOut dbg, int mode) {
  __shared__ float sT[12];
  __shared__ LmTrig sTrig;
  __shared__ float rows[S2M_THREADS][8];   // 6 Jacobian entries, rhs, accepted flag
  __shared__ double red[S2M_THREADS / 32][S2M_SUMS];
  __shared__ int s_ties;
  __shared__ bool s_last;

  if (mode == 0 && st->done) return;
  const int tid = threadIdx.x;
  if (tid == 0) {
    if (T_override) {
      for (int k = 0; k < 12; ++k) sT[k] = T_override[k];
    } else {
      pose_to_T(st->pose, sT);  // updatePointAssociateToMap (:1613-1616)
    }
    sTrig = lm_trig(st->pose);
    s_ties = 0;
  }
  __syncthreads();

  const int i = blockIdx.x * S2M_THREADS + tid;
  float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float rhs = 0.f;
  bool flag = false, tie = false;
  if (i < nq) {
    const float4 ori = scan[i];
    const float4 sel = apply_T(sT, ori);
    Top5 t;
    grid_knn5(sel, g, map_sorted, cell_start, t);
    float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.d4 < g.gate_d2) {  // :1641 (the gate is the reference's 1.0)
      float4 nbr[5];
      nbr[0] = __ldg(map4 + t.i0); nbr[1] = __ldg(map4 + t.i1); nbr[2] = __ldg(map4 + t.i2);
      nbr[3] = __ldg(map4 + t.i3); nbr[4] = __ldg(map4 + t.i4);
      flag = plane_residual(ori, sel, nbr, coeff);
      tie = (t.d0 == t.d1) || (t.d1 == t.d2) || (t.d2 == t.d3) || (t.d3 == t.d4) || (t.d4 == t.rej);
      if (!flag) coeff = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (flag) jacobian_row(sTrig, ori, coeff, row, rhs);
    if (mode == 1) {
      if (dbg.nn_idx) {
        int* o = dbg.nn_idx + (size_t)i * 5;
        o[0] = t.i0; o[1] = t.i1; o[2] = t.i2; o[3] = t.i3; o[4] = t.i4;
      }
      if (dbg.nn_d2) {
        float* o = dbg.nn_d2 + (size_t)i * 5;
        o[0] = t.d0; o[1] = t.d1; o[2] = t.d2; o[3] = t.d3; o[4] = t.d4;
      }
      if (dbg.coeff) dbg.coeff[i] = coeff;
      if (dbg.flag) dbg.flag[i] = flag ? 1 : 0;
      if (dbg.tie) dbg.tie[i] = tie ? 1 : 0;
    }
  }
  if (mode == 1) return;

#pragma unroll
  for (int k = 0; k < 6; ++k) rows[tid][k] = row[k];
  rows[tid][6] = rhs;
  rows[tid][7] = flag ? 1.f : 0.f;
  if (flag && tie) atomicAdd(&s_ties, 1);
  __syncthreads();

  // 27 FP64 sums + count: thread (slice, p) adds its 32 rows' product p; products of two floats are
  // exact in double, so only the order of additions differs from cv::gemm's.
  {
    const int p = tid & 31, slice = tid >> 5;
    int a = 0, b = 0;
    if (p < 21) {  // upper-triangle pair index -> (a, b)
      int q = p;
      a = 0;
      while (q >= 6 - a) { q -= 6 - a; ++a; }
      b = a + q;
    } else if (p < 27) {
      a = p - 21; b = 6;
    } else {
      a = 7; b = 7;  // p == 27: accepted count (flag*flag); p > 27 unused (adds zeros below)
    }
    double acc = 0.0;
    if (p <= 27) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float* rr = rows[slice * 32 + r];
        acc += (double)rr[a] * (double)rr[b];
      }
    }
    red[slice][p] = acc;
  }
  __syncthreads();
  if (tid < S2M_SUMS) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < S2M_THREADS / 32; ++k) s += red[k][tid];
    if (tid == 28) s = (double)s_ties;
    partials[(size_t)blockIdx.x * S2M_SUMS + tid] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {  // fixed-order sum over blocks: 8 slices of blocks, then the slices
    const int p = tid & 31, slice = tid >> 5;
    double acc = 0.0;
    for (unsigned b = slice; b < gridDim.x; b += S2M_THREADS / 32) acc += __ldcg(partials + (size_t)b * S2M_SUMS + p);
    red[slice][p] = acc;
  }
  __syncthreads();
  if (tid < S2M_SUMS) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < S2M_THREADS / 32; ++k) s += red[k][tid];
    red[0][tid] = s;
  }
  __syncthreads();
  if (tid == 0) {
    lm_finalize(st, red[0]);
    *ticket = 0u;
  }
}

static int check_grid(Ctx* c) {
  if (!c->grid_valid) { c->err = "no local map installed (call liogpu_set_local_map / liogpu_build_local_map)"; return LIOGPU_E_NO_MAP; }
  return LIOGPU_OK;
}

int scan2map_dev(Ctx* c, const float4* scan4, int n, float pose_io[6], float matP_io[36], int* degenerate_io,
                 int max_iter, liogpu_s2m_info* info) {
  int rc = check_grid(c);
  if (rc != LIOGPU_OK) return rc;
  if (max_iter < 1 || max_iter > LIOGPU_MAX_ITER) { c->err = "max_iter out of range"; return LIOGPU_E_INVALID; }
  const int blocks = div_up(n, S2M_THREADS);
  LIOGPU_CUDA_OK(c, c->lm_state.reserve(sizeof(LmDevState)));
  LIOGPU_CUDA_OK(c, c->partials.reserve((size_t)blocks * S2M_SUMS * sizeof(double)));
  if (!c->block_counter.p) {
    LIOGPU_CUDA_OK(c, c->block_counter.reserve(64));
    LIOGPU_CUDA_OK(c, cudaMemsetAsync(c->block_counter.p, 0, 64, c->stream));
  }
  LmDevState* h = reinterpret_cast<LmDevState*>((char*)c->h_pinned + 4096);
  memset(h, 0, sizeof(LmDevState));
  for (int k = 0; k < 6; ++k) h->pose[k] = pose_io[k];
  for (int k = 0; k < 36; ++k) h->matP[k] = matP_io[k];
  h->degenerate = *degenerate_io;
  h->max_iter = max_iter;
  LmDevState* d = c->lm_state.as<LmDevState>();
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(d, h, sizeof(LmDevState), cudaMemcpyHostToDevice, c->stream));
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
  SurfDebugOut dbg{nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int it = 0; it < max_iter; ++it) {
    s2m_iter_kernel<<<blocks, S2M_THREADS, 0, c->stream>>>(scan4, n, c->map4.as<float4>(), c->map_sorted.as<float4>(),
                                                           c->cell_start.as<uint32_t>(), c->grid, d,
                                                           c->partials.as<double>(), c->block_counter.as<unsigned>(),
                                                           nullptr, dbg, 0);
  }
  c->launches += max_iter;
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h, d, sizeof(LmDevState), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  LIOGPU_CUDA_OK(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  for (int k = 0; k < 6; ++k) pose_io[k] = h->pose[k];
  for (int k = 0; k < 36; ++k) matP_io[k] = h->matP[k];
  *degenerate_io = h->degenerate;
  if (info) {
    info->iterations = h->iter;
    info->converged = h->converged;
    info->n_query = n;
    info->n_sel = h->n_sel;
    info->is_degenerate = h->degenerate;
    info->tie_queries = h->tie_queries;
    info->delta_r_deg = h->delta_r;
    info->delta_t_cm = h->delta_t;
    memcpy(info->JtJ, h->JtJ, sizeof(info->JtJ));
    memcpy(info->Jtr, h->Jtr, sizeof(info->Jtr));
    memcpy(info->pose_hist, h->pose_hist, sizeof(info->pose_hist));
    memcpy(info->nsel_hist, h->nsel_hist, sizeof(info->nsel_hist));
    info->gpu_ms = c->last_ms;
  }
  return LIOGPU_OK;
}

int surf_optimization_dev(Ctx* c, const float4* scan4, int n, const float* pose6, const float* T12, int* nn_idx,
                          float* nn_d2, float* coeff, unsigned char* flag, unsigned char* tie) {
  int rc = check_grid(c);
  if (rc != LIOGPU_OK) return rc;
  if ((pose6 == nullptr) == (T12 == nullptr)) { c->err = "exactly one of pose6 / T12 must be given"; return LIOGPU_E_INVALID; }
  if (n <= 0) return LIOGPU_OK;
  LIOGPU_CUDA_OK(c, c->lm_state.reserve(sizeof(LmDevState)));
  LIOGPU_CUDA_OK(c, c->partials.reserve((size_t)div_up(n, S2M_THREADS) * S2M_SUMS * sizeof(double)));
  LIOGPU_CUDA_OK(c, c->misc.reserve(256));
  LIOGPU_CUDA_OK(c, c->dbg_idx.reserve((size_t)n * 5 * sizeof(int)));
  LIOGPU_CUDA_OK(c, c->dbg_d2.reserve((size_t)n * 5 * sizeof(float)));
  LIOGPU_CUDA_OK(c, c->dbg_coeff.reserve((size_t)n * sizeof(float4)));
  LIOGPU_CUDA_OK(c, c->dbg_flag.reserve((size_t)n));
  LIOGPU_CUDA_OK(c, c->dbg_tie.reserve((size_t)n));
  if (!c->block_counter.p) {
    LIOGPU_CUDA_OK(c, c->block_counter.reserve(64));
    LIOGPU_CUDA_OK(c, cudaMemsetAsync(c->block_counter.p, 0, 64, c->stream));
  }
  LmDevState* h = reinterpret_cast<LmDevState*>((char*)c->h_pinned + 4096);
  memset(h, 0, sizeof(LmDevState));
  if (pose6) for (int k = 0; k < 6; ++k) h->pose[k] = pose6[k];
  LmDevState* d = c->lm_state.as<LmDevState>();
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(d, h, sizeof(LmDevState), cudaMemcpyHostToDevice, c->stream));
  float* d_T = nullptr;
  if (T12) {
    d_T = c->misc.as<float>() + 16;
    float* hT = reinterpret_cast<float*>((char*)c->h_pinned + 3072);
    for (int k = 0; k < 12; ++k) hT[k] = T12[k];
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(d_T, hT, 12 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  SurfDebugOut dbg{c->dbg_idx.as<int>(), c->dbg_d2.as<float>(), c->dbg_coeff.as<float4>(),
                   c->dbg_flag.as<unsigned char>(), c->dbg_tie.as<unsigned char>()};
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev0, c->stream));
  s2m_iter_kernel<<<div_up(n, S2M_THREADS), S2M_THREADS, 0, c->stream>>>(
      scan4, n, c->map4.as<float4>(), c->map_sorted.as<float4>(), c->cell_start.as<uint32_t>(), c->grid, d,
      c->partials.as<double>(), c->block_counter.as<unsigned>(), d_T, dbg, 1);
  c->launches++;
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, cudaEventRecord(c->ev1, c->stream));
  if (nn_idx) LIOGPU_CUDA_OK(c, cudaMemcpyAsync(nn_idx, dbg.nn_idx, (size_t)n * 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (nn_d2) LIOGPU_CUDA_OK(c, cudaMemcpyAsync(nn_d2, dbg.nn_d2, (size_t)n * 5 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (coeff) LIOGPU_CUDA_OK(c, cudaMemcpyAsync(coeff, dbg.coeff, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (flag) LIOGPU_CUDA_OK(c, cudaMemcpyAsync(flag, dbg.flag, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (tie) LIOGPU_CUDA_OK(c, cudaMemcpyAsync(tie, dbg.tie, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  LIOGPU_CUDA_OK(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  return LIOGPU_OK;
}

}  // namespace liogpu
