// s2m_fused.cuh — the whole loop of scan2MapOptimization (mapOptmization.cpp:1848-1859) as ONE persistent
// cooperative launch per registration.  Included by s2m.cu (same translation unit: it reuses the search, plane-fit
// and 6x6 routines defined there).
//
// Grid: every CTA slot of the GPU (occupancy x SM count), co-resident by cooperative launch.  CTA `gridDim.x - 1`
// is a service CTA (iteration 0's eigen-decomposition / matP, off the critical path); the others loop:
//
//   per Gauss-Newton iteration
//     main phase   chunks of 256 sweep points pulled from an atomic queue (no wave tail: a CTA that finishes
//                  early takes the next chunk).  Per chunk, three steps separated by __syncthreads:
//                    1 classify  thread = point.  From iteration 1 on every point carries a CANDIDATE SET: the (up
//                                to) 8 nearest map points found by its last search and a lower bound `lb` on the
//                                distance from the point to every map point outside the set.  The set members'
//                                distances to the moved point are evaluated exactly; if the 5th smallest is
//                                below lb - |move| (triangle inequality, rounding margins included) the five
//                                nearest neighbours of the moved point are certainly inside the set: they are
//                                selected and ordered by (d2, map index) with no grid walk at all — an exact
//                                certificate, not a heuristic.  Otherwise the point is queued for a search.
//                    2 search    the queued points, COMPACTED over the CTA (dense lanes), walk the sorted grid
//                                for the 9 nearest inside the seeded bound (or the phase-1 gate): exact 5-NN +
//                                the next candidate set + its bound.  Points phase 1 cannot settle go to the
//                                chunk's segment of the leftover list.
//                    3 fit       thread = point: 5x3 plane fit, weight, Jacobian row; FP64 block reduction of the
//                                27 sums of A^T A / A^T b into the chunk's partial row.
//                  A chunk with at most 8 such points finishes them on the spot (one warp-cooperative full-gate search
//                  per warp); a chunk with more defers them to the grid-wide leftover phase.
//     ticket A     the last CTA to finish the main phase looks at the number of deferred leftovers:
//                    none (the usual case from iteration 1 on): it adds the chunk rows in a fixed order, runs the
//                          6x6 tail of LMOptimization (lm_finalize_warp) and releases the others — ONE grid-wide
//                          synchronisation per iteration;
//                    some: it opens the leftover phase: warp-cooperative full-gate search, one point per warp over
//                          the whole grid (static map -> fixed summation order), per-CTA partial rows, ticket B, the
//                          last CTA adds chunk rows + CTA rows, runs the tail and releases the others.
//
// Nothing returns to the host inside the loop; the 1.8 KB state block is read back once.
#pragma once

namespace liogpu {

constexpr int FZ_THREADS = 256;
constexpr int FZ_WARPS = FZ_THREADS / 32;
#ifndef FZ_MINBLOCKS_CFG
#define FZ_MINBLOCKS_CFG 3
#endif
constexpr int FZ_K = LIOGPU_FZ_K;        // members of a candidate set
constexpr int FZ_MAXCHUNKS = 4096;       // chunk-offset table of the leftover phase lives in shared memory
constexpr float FZ_REL = 1e-5f;          // relative safety margin of every bound (f32 rounding is < 3e-7)
constexpr float FZ_SEED_MARGIN = 0.10f;  // a seeded search enumerates this far (m) beyond the seeds' 5th distance
constexpr int FZ_DIRECT_MIN = 128;       // chunks with more search requests than this skip the compaction (thread = point)
constexpr int FZ_INPLACE = FZ_WARPS;     // chunks with at most this many leftovers finish them on the spot
constexpr int FZ_PROBES = 8;             // %globaltimer stamps per iteration (profile_kernels)

// ---- 9 best (d2, map index) pairs, same 64-bit keys as Top5; slot FZ_K is the pruning threshold ----
struct TopN {
  u64 k[FZ_K + 1];
  __device__ __forceinline__ void init(float gate) {
#pragma unroll
    for (int j = 0; j <= FZ_K; ++j) k[j] = ((u64)__float_as_uint(gate)) << 32;
  }
  __device__ __forceinline__ void offer(float d, int id) {
    const u64 key = (((u64)__float_as_uint(d)) << 32) | (u64)(unsigned)id;
    if (key < k[FZ_K]) {
      k[FZ_K] = key;
#pragma unroll
      for (int j = FZ_K; j > 0; --j) {
        const u64 lo = min(k[j - 1], k[j]), hi = max(k[j - 1], k[j]);
        k[j - 1] = lo; k[j] = hi;
      }
    }
  }
  __device__ __forceinline__ float d(int j) const { return __uint_as_float((unsigned)(k[j] >> 32)); }
  __device__ __forceinline__ int i(int j) const { return (int)(unsigned)(k[j] & 0xffffffffull); }
  __device__ __forceinline__ float worst() const { return d(FZ_K); }
};

// grid_knn5 with the wider list: on return every map point that is NOT in k[0..FZ_K-1] has d2 >= worst(), and
// k[0..4] are the exact 5 nearest among the points closer than sqrt(gate_d2) (ascending (d2, index)).
__device__ __forceinline__ void grid_knn_topn(const float4 q, const GridParams& g, const float gate_d2,
                                              const float4* __restrict__ map_sorted,
                                              const uint32_t* __restrict__ cell_start, TopN& t) {
  t.init(gate_d2);
  const float s2 = 2.0f * g.slack;
  const float reach = sqrtf(gate_d2) * 1.000001f + s2;
  int zmin = (int)floorf((q.z - reach - g.oz) * g.inv_h), zmax = (int)floorf((q.z + reach - g.oz) * g.inv_h);
  int ymin = (int)floorf((q.y - reach - g.oy) * g.inv_h), ymax = (int)floorf((q.y + reach - g.oy) * g.inv_h);
  zmin = max(zmin, 0); zmax = min(zmax, g.nz - 1);
  ymin = max(ymin, 0); ymax = min(ymax, g.ny - 1);
  if (zmin > zmax || ymin > ymax) return;
  const int cz = min(max((int)floorf((q.z - g.oz) * g.inv_h), zmin), zmax);
  const int cy = min(max((int)floorf((q.y - g.oy) * g.inv_h), ymin), ymax);
  const int nzs = zmax - zmin + 1, nys = ymax - ymin + 1;
  for (int kz = 0, seen_z = 0; seen_z < nzs; ++kz) {
    const int z = cz + zigzag(kz);
    if (z < zmin || z > zmax) continue;
    ++seen_z;
    const float zlo = g.oz + (float)z * g.h;
    const float gz = fmaxf(fmaxf(zlo - q.z, q.z - (zlo + g.h)) - s2, 0.f);
    const float gz2 = gz * gz;
    if (gz2 > t.worst()) continue;
    for (int ky = 0, seen_y = 0; seen_y < nys; ++ky) {
      const int y = cy + zigzag(ky);
      if (y < ymin || y > ymax) continue;
      ++seen_y;
      const float ylo = g.oy + (float)y * g.h;
      const float gy = fmaxf(fmaxf(ylo - q.y, q.y - (ylo + g.h)) - s2, 0.f);
      const float m2 = (gz2 + gy * gy) * 0.999999f;
      const float worst = t.worst();
      if (m2 > worst) continue;
      const float r = sqrtf(worst - m2) * 1.000001f + s2;
      int xlo = (int)floorf((q.x - r - g.ox) * g.inv_h);
      int xhi = (int)floorf((q.x + r - g.ox) * g.inv_h);
      xlo = max(xlo, 0);
      xhi = min(xhi, g.nx - 1);
      if (xlo > xhi) continue;
      const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
      const uint32_t s = __ldg(cell_start + row + xlo);
      const uint32_t e = __ldg(cell_start + row + xhi + 1);
      for (uint32_t j = s; j < e; j += 4) {
        const uint32_t last = e - 1;
        const float4 p0 = __ldg(map_sorted + j);
        const float4 p1 = __ldg(map_sorted + min(j + 1, last));
        const float4 p2 = __ldg(map_sorted + min(j + 2, last));
        const float4 p3 = __ldg(map_sorted + min(j + 3, last));
        t.offer(l2_simple(q, p0), __float_as_int(p0.w));
        if (j + 1 < e) t.offer(l2_simple(q, p1), __float_as_int(p1.w));
        if (j + 2 < e) t.offer(l2_simple(q, p2), __float_as_int(p2.w));
        if (j + 3 < e) t.offer(l2_simple(q, p3), __float_as_int(p3.w));
      }
    }
  }
}

#define FZ_CE(a, b) { const u64 lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
// 19-comparator sorting network for 8 keys (Batcher odd-even merge sort), ascending
__device__ __forceinline__ void sort8(u64 k[8]) {
  FZ_CE(k[0], k[1]); FZ_CE(k[2], k[3]); FZ_CE(k[4], k[5]); FZ_CE(k[6], k[7]);
  FZ_CE(k[0], k[2]); FZ_CE(k[1], k[3]); FZ_CE(k[4], k[6]); FZ_CE(k[5], k[7]);
  FZ_CE(k[1], k[2]); FZ_CE(k[5], k[6]);
  FZ_CE(k[0], k[4]); FZ_CE(k[1], k[5]); FZ_CE(k[2], k[6]); FZ_CE(k[3], k[7]);
  FZ_CE(k[2], k[4]); FZ_CE(k[3], k[5]);
  FZ_CE(k[1], k[2]); FZ_CE(k[3], k[4]); FZ_CE(k[5], k[6]);
}

struct FusedArgs {
  const float4* scan;
  int nq;
  const float4* map4;
  const float4* map_sorted;
  const uint32_t* cell_start;
  GridParams g;
  LmDevState* st;
  double* chunk_rows;      // [nchunks][S2M_SUMS]  partial sums of a chunk (deterministic: a chunk is a fixed set of points)
  double* cta_rows;        // [compute CTAs][S2M_SUMS]
  int* left_list;          // [nchunks][FZ_THREADS] leftover point indices, per chunk segment
  int* chunk_nleft;        // [nchunks]
  int* prev_nn;            // [FZ_K][nq] candidate set of every point (-1: empty slot; prev_nn[0][i] < 0: no set)
  float* prev_lb;          // [nq] lower bound on the distance to every map point outside the set
  float4* hopeless;        // [nq] see HOPELESS_MARGIN
  unsigned long long* probe;  // [LIOGPU_MAX_ITER][FZ_PROBES] %globaltimer stamps (profile_kernels), or null
  SurfDebugOut dbg;        // per-point outputs of the last executed iteration (trace entry point), or nulls
  int nchunks;
  int use_cert;            // 0: never take the certificate (A/B switch: every point with a set is searched)
};

__device__ __forceinline__ unsigned long long fz_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned fz_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}


// transPointAssociateToMap + LM trig of `pose` into caller storage (same arithmetic as warp_refresh_transform)
__device__ __forceinline__ void fz_warp_transform(const int lane, const float pose_l, float* T, float* trig) {
  float sn = 0.f, cs = 0.f;
  if (lane < 3) {
    const double a = (double)pose_l;
    sn = (float)sin(a);
    cs = (float)cos(a);
  }
  const float F = __shfl_sync(FULL, sn, 0), E = __shfl_sync(FULL, cs, 0);
  const float D = __shfl_sync(FULL, sn, 1), C = __shfl_sync(FULL, cs, 1);
  const float B = __shfl_sync(FULL, sn, 2), A = __shfl_sync(FULL, cs, 2);
  const float px = __shfl_sync(FULL, pose_l, 3), py = __shfl_sync(FULL, pose_l, 4), pz = __shfl_sync(FULL, pose_l, 5);
  if (lane == 0) {
    const float DE = D * E, DF = D * F;
    T[0] = A * C;  T[1] = A * DF - B * E;  T[2] = B * F + A * DE;  T[3] = px;
    T[4] = B * C;  T[5] = A * E + B * DF;  T[6] = B * DE - A * F;  T[7] = py;
    T[8] = -D;     T[9] = C * F;           T[10] = C * E;          T[11] = pz;
    trig[0] = B; trig[1] = A; trig[2] = D; trig[3] = C; trig[4] = F; trig[5] = E;
  }
}

// iteration 0's matP when the certificate let the loop go ahead (body of lm_matp_kernel), one warp
__device__ __forceinline__ void fz_matp_warp(LmDevState* st, FinSmem& m, const int lane) {
  for (int e = lane; e < 36; e += 32) m.AtA[e] = __ldcg(st->AtA0 + e);
  __syncwarp();
  warp_eigen6(m, lane);
  for (int e = lane; e < 36; e += 32) m.V2[e] = m.V[e];
  __syncwarp();
  int deg = 0;
  if (lane == 0) {
    for (int i = 5; i >= 0; --i) {
      if (m.W[i] < 100.f) {
        for (int j = 0; j < 6; ++j) m.V2[i * 6 + j] = 0.f;
        deg = 1;
      } else {
        break;
      }
    }
  }
  __syncwarp();
  warp_inv6_lu(m, lane);
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e % 6;
    double acc = 0;
    for (int k = 0; k < 6; ++k) acc += (double)m.Vi[i * 6 + k] * (double)m.V2[k * 6 + j];
    st->matP[e] = (float)acc;
  }
  if (lane == 0) {
    if (deg) st->cert_mismatch = 1;
    st->eig_pending = 0;
  }
}

// shared-memory carve-up: the main phase and the leftover phase never overlap in time
struct FzMainSmem {
  int res_id[5][FZ_THREADS];      // the five neighbours of every slot of the chunk
  float bound[FZ_THREADS];        // search request: > 0 seeded bound, 0 phase-1 gate, -1 straight to the leftovers, -2 none
  float rows[FZ_THREADS][8];      // Jacobian row, rhs, accepted flag
  unsigned char res_meta[FZ_THREADS];  // bit0 found, bit1 tie
  unsigned char list[FZ_THREADS];      // slots queued for the search step, slot order
  unsigned char llist[FZ_THREADS];     // slots the search could not settle (leftovers), list order
};
struct FzLeftSmem {
  int off[FZ_MAXCHUNKS + 1];
  float rows[FZ_THREADS][8];
};
union FzSmem {
  FzMainSmem m;
  FzLeftSmem l;
};

// what a leftover search leaves behind for the next iteration: the candidate set (its five neighbours), the bound
// (everything else it visited was >= rej away, everything it did not visit is beyond the extended gate) and the
// hopeless marker
__device__ __forceinline__ void fz_store_leftover(const FusedArgs& A, const int mine, const float4 sel, const Top5& t,
                                                  const int n_ext, const float ge2, const bool found) {
  A.prev_nn[mine] = found ? t.i(t.k0) : -1;
  A.prev_nn[(size_t)A.nq + mine] = t.i(t.k1);
  A.prev_nn[2 * (size_t)A.nq + mine] = t.i(t.k2);
  A.prev_nn[3 * (size_t)A.nq + mine] = t.i(t.k3);
  A.prev_nn[4 * (size_t)A.nq + mine] = t.i(t.k4);
#pragma unroll
  for (int j = 5; j < FZ_K; ++j) A.prev_nn[(size_t)j * A.nq + mine] = -1;
  A.prev_lb[mine] = sqrtf(fminf(t.rej, ge2)) * (1.f - FZ_REL);
  const bool hopeless = !found && n_ext < 5;
  A.hopeless[mine] = make_float4(sel.x, sel.y, sel.z, hopeless ? 1.f : 0.f);
}

// sum of `nrows` partial rows (S2M_SUMS doubles each) in a fixed order: warp w takes rows w, w+8, ... with eight
// independent accumulators (eight 256-byte loads in flight per warp), then the warps are added in order.
// Result: lane l of warp 0 returns sum[l]; every thread must call it.
__device__ __forceinline__ double fz_reduce_rows(const double* __restrict__ rows, const int nrows, double (*red)[S2M_SUMS],
                                                 const int warp, const int lane) {
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  constexpr int W = FZ_WARPS;
  int b = warp;
  for (; b + 7 * W < nrows; b += 8 * W) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldcg(rows + (size_t)(b + k * W) * S2M_SUMS + lane);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += v[k];
  }
  for (int k = 0; b < nrows; b += W, ++k) a[k] += __ldcg(rows + (size_t)b * S2M_SUMS + lane);
  __syncthreads();
  red[warp][lane] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  double sum = 0.0;
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < FZ_WARPS; ++k) sum += red[k][lane];
  }
  return sum;
}

__global__ void __launch_bounds__(FZ_THREADS, FZ_MINBLOCKS_CFG)
s2m_fused_kernel(const FusedArgs A) {
  __shared__ float sT[12], sTp[12];
  __shared__ LmTrig sTrig;
  __shared__ __align__(16) FzSmem sm;
  __shared__ double red[FZ_WARPS][S2M_SUMS];
  __shared__ double s_sum[S2M_SUMS];
  __shared__ FinSmem s_fin;
  __shared__ int s_chunk, s_wcnt[FZ_WARPS], s_wcnt2[FZ_WARPS], s_misc[4];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LmDevState* const st = A.st;
  const int G = (int)gridDim.x - 1;  // compute CTAs

  if ((int)blockIdx.x == G) {
    // ---- service CTA: iteration 0's eigen-decomposition + matP while iterations 1.. run ----
    if (tid < 32) {
      if (lane == 0) { while (fz_ld_acquire(&st->fz_phase) < 2u) { } __threadfence(); }
      __syncwarp();
      if (__ldcg(&st->eig_pending)) fz_matp_warp(st, s_fin, lane);
    }
    return;
  }

  const bool can_phase1 = A.g.gate1_d2 < A.g.gate_d2;  // dense map: cheap first phase inside a small gate
  const int max_iter = st->max_iter;
  const RowAcc ra = row_acc_of(lane);
  const float ge = sqrtf(A.g.gate_d2) + HOPELESS_MARGIN;  // extended gate of the leftover search
  const float ge2 = ge * ge;

  for (int it = 0; it < max_iter; ++it) {
    unsigned long long* const probe = A.probe ? A.probe + it * FZ_PROBES : nullptr;
    // ---- this iteration's transform (updatePointAssociateToMap, :1613-1616) ----
    if (it == 0) {
      if (tid < 32) {
        const float pose_l = lane < 6 ? st->pose[lane] : 0.f;
        fz_warp_transform(lane, pose_l, sT, &sTrig.srx);
      }
    } else {
      if (tid < 12) { sT[tid] = __ldcg(st->T + tid); sTp[tid] = __ldcg(st->T_prev + tid); }
      if (tid == 32) {
        sTrig.srx = __ldcg(st->trig + 0); sTrig.crx = __ldcg(st->trig + 1); sTrig.sry = __ldcg(st->trig + 2);
        sTrig.cry = __ldcg(st->trig + 3); sTrig.srz = __ldcg(st->trig + 4); sTrig.crz = __ldcg(st->trig + 5);
      }
    }
    __syncthreads();
    if (probe && blockIdx.x == 0 && tid == 0) probe[0] = fz_globaltimer();
    int cta_deferred = 0;  // meaningful in thread 0

    // =========================== main phase: dynamic queue of 256-point chunks ===========================
    for (;;) {
      if (tid == 0) s_chunk = (int)atomicAdd(&st->fz_queue[it], 1u);
      __syncthreads();
      const int c = s_chunk;
      if (c >= A.nchunks) break;
      const int base = c * FZ_THREADS;
      const int i = base + tid;
      // ---- step 1: classify ----
      float req = -2.f;
      int n_seeded = 0, n_cert = 0;
      unsigned char meta = 0;
      if (i < A.nq) {
        const float4 ori = A.scan[i];
        const float4 sel = apply_T(sT, ori);
        req = can_phase1 ? 0.f : -1.f;
        if (it > 0) {
          const int p0 = __ldcg(A.prev_nn + i);
          if (p0 < 0) {
            const float4 hr = __ldcg(A.hopeless + i);
            if (hr.w > 0.f) {
              const float dx = sel.x - hr.x, dy = sel.y - hr.y, dz = sel.z - hr.z;
              const float lim = HOPELESS_MARGIN - 1e-3f;
              if ((dx * dx + dy * dy + dz * dz) < lim * lim) req = -2.f;  // still cannot have 5 neighbours within the gate
            }
          } else {
            u64 key[8];
            int id[8];
            id[0] = p0;
#pragma unroll
            for (int j = 1; j < FZ_K; ++j) id[j] = __ldcg(A.prev_nn + (size_t)j * A.nq + i);
#pragma unroll
            for (int j = 0; j < FZ_K; ++j) {
              key[j] = ~0ull;
              if (id[j] >= 0) {
                const float d = l2_simple(sel, __ldg(A.map4 + id[j]));
                key[j] = (((u64)__float_as_uint(d)) << 32) | (u64)(unsigned)id[j];
              }
            }
            sort8(key);
            const float D5 = __uint_as_float((unsigned)(key[4] >> 32));
            const float bound = __uint_as_float(__float_as_uint(D5) + 1u);  // next float above: the seeds stay inside
            if (bound <= A.g.gate_d2) {
              n_seeded = 1;
              req = bound;
              // certificate: every map point outside the set was >= lb away from where this point stood when the
              // set was built (or last certified); it has moved by |sel - prev|
              const float4 pv = apply_T(sTp, ori);
              const float move = sqrtf(l2_simple(sel, pv));
              const float L = __ldcg(A.prev_lb + i) - move * (1.f + FZ_REL) - 1e-7f;
              if (A.use_cert && L > 0.f && L * L * (1.f - FZ_REL) > D5) {
                n_cert = 1;
                req = -2.f;
#pragma unroll
                for (int j = 0; j < 5; ++j) sm.m.res_id[j][tid] = (int)(unsigned)(key[j] & 0xffffffffull);
                const float d0 = __uint_as_float((unsigned)(key[0] >> 32)), d1 = __uint_as_float((unsigned)(key[1] >> 32));
                const float d2 = __uint_as_float((unsigned)(key[2] >> 32)), d3 = __uint_as_float((unsigned)(key[3] >> 32));
                const float d5 = __uint_as_float((unsigned)(key[5] >> 32));
                const bool tie = d0 == d1 || d1 == d2 || d2 == d3 || d3 == D5 || (key[5] != ~0ull && D5 == d5);
                meta = (unsigned char)(1 | (tie ? 2 : 0));
                A.prev_lb[i] = L * (1.f - FZ_REL);
              }
            }
          }
        }
      }
      sm.m.res_meta[tid] = meta;
      sm.m.bound[tid] = req;
      const bool want = req > -2.f;
      const unsigned wm = __ballot_sync(FULL, want);
      if (lane == 0) s_wcnt[warp] = __popc(wm);
      __syncthreads();
      int ns = 0, woff = 0;
#pragma unroll
      for (int w = 0; w < FZ_WARPS; ++w) { if (w < warp) woff += s_wcnt[w]; ns += s_wcnt[w]; }
      // search-heavy chunk (iterations 0 and 1): thread = point, no compaction; otherwise the requests are
      // compacted so that the searching warps run with dense lanes
      const bool direct = ns > FZ_DIRECT_MIN;
      if (!direct) {
        if (want) sm.m.list[woff + __popc(wm & ((1u << lane) - 1u))] = (unsigned char)tid;
        __syncthreads();
      }
      // ---- step 2: search ----
      bool need2 = false;
      int slot = -1;
      if (direct) { if (want) slot = tid; }
      else if (tid < ns) slot = sm.m.list[tid];
      if (slot >= 0) {
        const int qi = base + slot;
        const float rq = sm.m.bound[slot];
        need2 = true;
        if (rq >= 0.f) {
          const float4 sel = apply_T(sT, A.scan[qi]);
          const bool seeded = rq > 0.f;
          float gate_use = A.g.gate1_d2;
          if (seeded) { const float r = sqrtf(rq) + FZ_SEED_MARGIN; gate_use = fmaxf(r * r, rq); }
          TopN t;
          grid_knn_topn(sel, A.g, gate_use, A.map_sorted, A.cell_start, t);
          const float lim5 = seeded ? rq : A.g.gate1_d2;
          need2 = !(t.d(4) < lim5);   // a seeded search always finds its five (the seeds are inside the bound)
          if (!need2) {
#pragma unroll
            for (int j = 0; j < 5; ++j) sm.m.res_id[j][slot] = t.i(j);
            const bool tie = t.d(0) == t.d(1) || t.d(1) == t.d(2) || t.d(2) == t.d(3) || t.d(3) == t.d(4) || t.d(4) == t.d(5);
            sm.m.res_meta[slot] = (unsigned char)(1 | (tie ? 2 : 0));
            // next iteration's candidate set and its bound: slots still holding the sentinel are empty
            const unsigned gbits = __float_as_uint(gate_use);
#pragma unroll
            for (int j = 0; j < FZ_K; ++j) {
              const bool real = !((unsigned)(t.k[j] >> 32) == gbits && (unsigned)t.k[j] == 0u);
              A.prev_nn[(size_t)j * A.nq + qi] = real ? t.i(j) : -1;
            }
            A.prev_lb[qi] = sqrtf(t.worst()) * (1.f - FZ_REL);
          }
        }
      }
      // leftovers of this chunk, in search order
      int nl = 0;
      {
        const unsigned fm = __ballot_sync(FULL, need2);
        if (lane == 0) s_wcnt2[warp] = __popc(fm);
        __syncthreads();
        int off = 0;
#pragma unroll
        for (int w = 0; w < FZ_WARPS; ++w) { if (w < warp) off += s_wcnt2[w]; nl += s_wcnt2[w]; }
        if (need2) sm.m.llist[off + __popc(fm & ((1u << lane) - 1u))] = (unsigned char)slot;
      }
      const bool deferred = nl > FZ_INPLACE;
      if (nl > 0) {
        __syncthreads();
        if (!deferred) {
          // a handful: one warp-cooperative full-gate search per warp, right here
          if (warp < nl) {
            const int ls = sm.m.llist[warp];
            const int mine = base + ls;
            const float4 sel = apply_T(sT, A.scan[mine]);
            Top5 t;
            const int n_ext = warp_knn5(sel, A.g, ge2, A.map_sorted, A.cell_start, lane, t);
            if (lane == 0) {
              const bool found = t.d(t.k4) < A.g.gate_d2;  // :1641
              fz_store_leftover(A, mine, sel, t, n_ext, ge2, found);
              if (found) {
                sm.m.res_id[0][ls] = t.i(t.k0); sm.m.res_id[1][ls] = t.i(t.k1); sm.m.res_id[2][ls] = t.i(t.k2);
                sm.m.res_id[3][ls] = t.i(t.k3); sm.m.res_id[4][ls] = t.i(t.k4);
                sm.m.res_meta[ls] = (unsigned char)(1 | (t.tie() ? 2 : 0));
              }
            }
          }
        } else {
          if (tid < nl) A.left_list[(size_t)base + tid] = base + sm.m.llist[tid];
          if (tid == 0) cta_deferred += nl;
        }
      }
      if (tid == 0) A.chunk_nleft[c] = deferred ? nl : 0;
      __syncthreads();
      // ---- step 3: plane fit + Jacobian row (thread = point) ----
      float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float rhs = 0.f;
      bool flag = false, tie = false;
      if (i < A.nq) {
        const unsigned char mt = sm.m.res_meta[tid];
        float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        int nid[5] = {-1, -1, -1, -1, -1};
        float nd2[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        const bool pending = deferred && sm.m.bound[tid] > -2.f && !(mt & 1);  // finished by the leftover phase
        if (mt & 1) {
          const float4 ori = A.scan[i];
          const float4 sel = apply_T(sT, ori);
          float4 nbr[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) { nid[j] = sm.m.res_id[j][tid]; nbr[j] = __ldg(A.map4 + nid[j]); }
          flag = plane_residual(ori, sel, nbr, coeff);
          tie = (mt & 2) != 0;
          if (!flag) coeff = make_float4(0.f, 0.f, 0.f, 0.f);
          if (flag) jacobian_row(sTrig, ori, coeff, row, rhs);
          if (A.dbg.nn_d2) {
#pragma unroll
            for (int j = 0; j < 5; ++j) nd2[j] = l2_simple(sel, nbr[j]);
          }
        }
        if (!pending) {
          if (A.dbg.nn_idx) { int* o = A.dbg.nn_idx + (size_t)i * 5; for (int j = 0; j < 5; ++j) o[j] = nid[j]; }
          if (A.dbg.nn_d2) { float* o = A.dbg.nn_d2 + (size_t)i * 5; for (int j = 0; j < 5; ++j) o[j] = (mt & 1) ? nd2[j] : A.g.gate_d2; }
          if (A.dbg.coeff) A.dbg.coeff[i] = coeff;
          if (A.dbg.flag) A.dbg.flag[i] = flag ? 1 : 0;
          if (A.dbg.tie) A.dbg.tie[i] = tie ? 1 : 0;
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) sm.m.rows[tid][k] = row[k];
      sm.m.rows[tid][6] = rhs;
      sm.m.rows[tid][7] = flag ? 1.f : 0.f;
      const int w_ties = __popc(__ballot_sync(FULL, flag && tie));
      const int w_seed = __popc(__ballot_sync(FULL, n_seeded != 0));
      const int w_cert = __popc(__ballot_sync(FULL, n_cert != 0));
      __syncthreads();
      {
        double acc = 0.0;
        if (ra.live) {
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const float* rr = sm.m.rows[warp * 32 + r];
            acc += (double)rr[ra.a] * (double)rr[ra.b];
          }
        }
        if (lane == 28) acc = (double)w_ties;
        if (lane == 29) acc = (double)w_seed;
        if (lane == 30) acc = (double)w_cert;
        if (lane == 31) acc = warp == 0 ? (double)nl : 0.0;
        red[warp][lane] = acc;
      }
      __syncthreads();
      if (tid < S2M_SUMS) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < FZ_WARPS; ++k) sum += red[k][tid];
        A.chunk_rows[(size_t)c * S2M_SUMS + tid] = sum;
      }
      __syncthreads();
    }

    // =========================== ticket A: the last CTA out of the main phase decides ===========================
    if (probe && blockIdx.x == 0 && tid == 0) probe[1] = fz_globaltimer();
    if (tid == 0) {
      if (cta_deferred) atomicAdd(&st->fz_deferred[it], (unsigned)cta_deferred);
      __threadfence();
      s_last = (atomicAdd(&st->fz_ticket_a[it], 1u) == (unsigned)(G - 1));
    }
    __syncthreads();
    bool left_phase = false;
    if (s_last) {
      __threadfence();
      if (tid == 0) {
        s_misc[2] = (int)__ldcg(&st->fz_deferred[it]);
        if (probe) probe[2] = fz_globaltimer();
      }
      __syncthreads();
      left_phase = s_misc[2] > 0;
      if (left_phase) {
        if (tid == 0) { __threadfence(); atomicExch(&st->fz_phase, (unsigned)(2 * it + 1)); }
      } else {
        const double sum = fz_reduce_rows(A.chunk_rows, A.nchunks, red, warp, lane);
        if (tid < S2M_SUMS) s_sum[tid] = sum;
        if (probe && tid == 0) probe[6] = fz_globaltimer();
        __syncthreads();
        if (tid < 32) {
          if (tid < 12) st->T_prev[tid] = sT[tid];  // where the points stood in this iteration
          if (tid == 0) {
            st->certified = (int)s_sum[30]; st->leftovers = (int)s_sum[31];
            st->cert_hist[it] = (int)s_sum[30]; st->left_hist[it] = (int)s_sum[31]; st->seed_hist[it] = (int)s_sum[29];
          }
          __syncwarp();
          lm_finalize_warp(st, s_sum, s_fin, tid);
          __syncwarp();
          if (tid == 0) {
            if (probe) probe[7] = fz_globaltimer();
            __threadfence();
            atomicExch(&st->fz_phase, (unsigned)(2 * it + 2));
          }
        }
      }
    }
    // everybody: wait for the decision
    if (tid == 0) {
      unsigned ph;
      while ((ph = fz_ld_acquire(&st->fz_phase)) < (unsigned)(2 * it + 1)) { }
      __threadfence();
      s_misc[3] = (int)ph;
    }
    __syncthreads();

    if (s_misc[3] == 2 * it + 1) {
      // =========================== deferred-leftover phase (whole grid) ===========================
      {
        // exclusive scan of the per-chunk deferred counts (every CTA builds its own copy)
        constexpr int PER = FZ_MAXCHUNKS / FZ_THREADS;  // 16
        int cnt[PER];
        int local = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int cidx = tid * PER + k;
          cnt[k] = cidx < A.nchunks ? __ldcg(A.chunk_nleft + cidx) : 0;
          local += cnt[k];
        }
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += v;
        }
        if (lane == 31) s_wcnt[warp] = incl;
        __syncthreads();
        int run = incl - local;
        for (int w = 0; w < warp; ++w) run += s_wcnt[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int cidx = tid * PER + k;
          if (cidx < A.nchunks) sm.l.off[cidx] = run;
          run += cnt[k];
        }
        if (tid == FZ_THREADS - 1) s_misc[0] = run;
        __syncthreads();
        if (tid == 0) sm.l.off[A.nchunks] = s_misc[0];
        __syncthreads();
      }
      const int total = s_misc[0];
      const int warps_per_grid = G * FZ_WARPS;
      const int per_warp = min(32, max(1, (total + warps_per_grid - 1) / warps_per_grid));
      const int nbatch = (total + per_warp - 1) / per_warp;
      double acc = 0.0;
      int ties = 0;
      for (int r0 = 0; r0 < nbatch; r0 += warps_per_grid) {
        const int batch = r0 + (int)blockIdx.x * FZ_WARPS + warp;
        float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float rhs = 0.f;
        bool flag = false, tie = false;
        if (batch < nbatch) {
          const int e0 = batch * per_warp;
          const int cnt = min(per_warp, total - e0);
          int mine = -1;
          if (lane < cnt) {
            const int e = e0 + lane;
            int lo = 0, hi = A.nchunks;  // invariant: off[lo] <= e < off[hi]
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (sm.l.off[mid] <= e) lo = mid; else hi = mid;
            }
            mine = __ldcg(A.left_list + (size_t)lo * FZ_THREADS + (e - sm.l.off[lo]));
          }
          float4 ori = make_float4(0.f, 0.f, 0.f, 0.f), sel = ori;
          if (mine >= 0) { ori = A.scan[mine]; sel = apply_T(sT, ori); }
          Top5 t;
          t.init(A.g.gate_d2);
          int my_ext = 0;
          for (int j = 0; j < cnt; ++j) {  // the warp searches for point j; lane j keeps the answer
            float4 q;
            q.x = __shfl_sync(FULL, sel.x, j); q.y = __shfl_sync(FULL, sel.y, j);
            q.z = __shfl_sync(FULL, sel.z, j); q.w = 0.f;
            Top5 tj;
            const int n_ext = warp_knn5(q, A.g, ge2, A.map_sorted, A.cell_start, lane, tj);
            if (lane == j) { t = tj; my_ext = n_ext; }
          }
          if (mine >= 0) {
            const bool found = t.d(t.k4) < A.g.gate_d2;  // :1641
            float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 nbr[5];
            if (found) {
              nbr[0] = __ldg(A.map4 + t.i(t.k0)); nbr[1] = __ldg(A.map4 + t.i(t.k1)); nbr[2] = __ldg(A.map4 + t.i(t.k2));
              nbr[3] = __ldg(A.map4 + t.i(t.k3)); nbr[4] = __ldg(A.map4 + t.i(t.k4));
              flag = plane_residual(ori, sel, nbr, coeff);
              tie = t.tie();
              if (!flag) coeff = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (flag) jacobian_row(sTrig, ori, coeff, row, rhs);
            fz_store_leftover(A, mine, sel, t, my_ext, ge2, found);
            if (A.dbg.nn_idx) {
              int* o = A.dbg.nn_idx + (size_t)mine * 5;
              o[0] = found ? t.i(t.k0) : -1; o[1] = found ? t.i(t.k1) : -1; o[2] = found ? t.i(t.k2) : -1;
              o[3] = found ? t.i(t.k3) : -1; o[4] = found ? t.i(t.k4) : -1;
            }
            if (A.dbg.nn_d2) {
              float* o = A.dbg.nn_d2 + (size_t)mine * 5;
              o[0] = t.d(t.k0); o[1] = t.d(t.k1); o[2] = t.d(t.k2); o[3] = t.d(t.k3); o[4] = t.d(t.k4);
            }
            if (A.dbg.coeff) A.dbg.coeff[mine] = coeff;
            if (A.dbg.flag) A.dbg.flag[mine] = flag ? 1 : 0;
            if (A.dbg.tie) A.dbg.tie[mine] = tie ? 1 : 0;
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) sm.l.rows[tid][k] = row[k];
        sm.l.rows[tid][6] = rhs;
        sm.l.rows[tid][7] = flag ? 1.f : 0.f;
        if (flag && tie) ++ties;
        __syncwarp();
        if (ra.live) {  // every warp owns the slice of rows its own lanes staged
#pragma unroll 8
          for (int r = 0; r < 32; ++r) {
            const float* rr = sm.l.rows[warp * 32 + r];
            acc += (double)rr[ra.a] * (double)rr[ra.b];
          }
        }
        __syncwarp();
      }
      {
        int wt = ties;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wt += __shfl_xor_sync(FULL, wt, o);
        if (lane == 28) acc += (double)wt;
      }
      red[warp][lane] = acc;
      __syncthreads();
      if (tid < S2M_SUMS) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < FZ_WARPS; ++k) sum += red[k][tid];
        if (tid == 31 && blockIdx.x == 0) sum += (double)total;
        A.cta_rows[(size_t)blockIdx.x * S2M_SUMS + tid] = sum;
      }
      if (probe && blockIdx.x == 0 && tid == 0) probe[4] = fz_globaltimer();
      // ---- ticket B: the last CTA adds everything in a fixed order and runs the 6x6 tail ----
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last = (atomicAdd(&st->fz_ticket_b[it], 1u) == (unsigned)(G - 1));
      __syncthreads();
      if (s_last) {
        __threadfence();
        if (probe && tid == 0) probe[5] = fz_globaltimer();
        const double sum_a = fz_reduce_rows(A.chunk_rows, A.nchunks, red, warp, lane);
        const double sum_b = fz_reduce_rows(A.cta_rows, G, red, warp, lane);
        if (tid < S2M_SUMS) s_sum[tid] = sum_a + sum_b;
        if (probe && tid == 0) probe[6] = fz_globaltimer();
        __syncthreads();
        if (tid < 32) {
          if (tid < 12) st->T_prev[tid] = sT[tid];
          if (tid == 0) {
            st->certified = (int)s_sum[30]; st->leftovers = (int)s_sum[31];
            st->cert_hist[it] = (int)s_sum[30]; st->left_hist[it] = (int)s_sum[31]; st->seed_hist[it] = (int)s_sum[29];
          }
          __syncwarp();
          lm_finalize_warp(st, s_sum, s_fin, tid);
          __syncwarp();
          if (tid == 0) {
            if (probe) probe[7] = fz_globaltimer();
            __threadfence();
            atomicExch(&st->fz_phase, (unsigned)(2 * it + 2));
          }
        }
      }
      if (tid == 0) {
        while (fz_ld_acquire(&st->fz_phase) < (unsigned)(2 * it + 2)) { }
        __threadfence();
      }
      __syncthreads();
    }
    if (tid == 0) s_misc[1] = __ldcg(&st->done);
    __syncthreads();
    if (s_misc[1]) break;
  }
}

}  // namespace liogpu
