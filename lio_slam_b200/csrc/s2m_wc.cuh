// s2m_wc.cuh — s2m_main_wc_kernel: the search of surfOptimization (mapOptmization.cpp:1622-1641) with the candidate
// evaluation done WARP-COOPERATIVELY at full lane occupancy.  Included by s2m.cu (same translation unit).
//
// STATUS: A/B variant (LIOGPU_MAIN=wc), parity-green (tests/test_gpu_variants.py), MEASURED SLOWER than s2m_main_kernel on
// config 3: 100-108 us on a late iteration, 230-290 us on iterations 0-1 (68 us average for the default).  ncu
// (profiles/r02_ncu_full_s2m_main_wc_kernel.csv): 26 of 32 lanes active as intended, but the flat candidate list costs ~60
// instructions per candidate (segment tracking, owner look-ups, shared-memory atomics), so a launch executes 33-75 M warp
// instructions, no fewer than the per-thread walk's 36-40 M, at issue rate 0.38 and 3 CTAs per SM.
//
// Why it was built: s2m_main_kernel gives every sweep point a thread that walks its own rows of grid cells; ncu shows that kernel
// issue bound with 17 of 32 lanes active per instruction — the lanes of a warp visit different numbers of rows and
// candidates, and the top-5 insertion runs at 6 of 32 lanes.  Here the irregular part is flattened:
//
//   A  lane = point     transform, bound of the search (the previous iteration's five neighbours give the seeded bound,
//                       clamped to the phase-1 gate; iteration 0: the phase-1 gate), the <= 3 x 3 rows of cells the ball
//                       touches -> up to nine contiguous ranges ("segments") of map_sorted per point, all cell-table
//                       look-ups in flight together.  The warp's non-empty segments are compacted into shared memory
//                       and prefix-summed: the warp now owns one flat list of candidates.
//   B  lane = candidate every lane takes an equal slice of the flat list (one binary search to find its first segment,
//                       then a flat loop): distance to the segment's owner, and if it is inside the owner's bound the
//                       key (d2, map index) is appended to the owner's survivor list (shared-memory counter) and counted
//                       in the owner's 16-bin histogram of d2.  No top-5 insertion, no per-point pruning state: every
//                       lane executes the same instruction stream.
//      An owner with more survivors than its list holds (a loose bound: iteration 0, or a large pose step) picks from
//      its histogram the first bin limit with >= 5 candidates below it; a second pass over its segments collects exactly
//      those.  Exact: the bin index is a monotone function of d2, so nothing outside the collected set can beat or tie
//      a member of it.
//   C  lane = point     top 5 of the (typically 5-9) survivors under the total order (d2, map index) — the same set the
//                       sequential walk selects, the tie flag included (it depends only on the multiset of distances) —
//                       then plane fit, Jacobian row and the FP64 block reduction exactly as s2m_main_kernel.
//
// Points the scheme does not fit walk the grid themselves as before (grid_knn5): a ball that touches more than 3 x 3
// rows (only possible in a float corner case once the bound is clamped to the phase-1 gate), more than 60,000 candidates,
// or a histogram bin that alone overflows the list (thousands of coincident map points).  Points with fewer than five map
// points inside the phase-1 gate go to the leftover list exactly as in s2m_main_kernel; outputs, partial-sum layout and
// leftover segments are identical, so s2m_left_kernel does not know which main kernel ran.
#pragma once

namespace liogpu {

constexpr int WC_ROWS = 9;                 // rows of cells a ball of radius <= cell edge can touch: 3 x 3
constexpr int WC_SEGS = 32 * WC_ROWS;      // segments per warp
#ifndef WC_CAP_CFG
#define WC_CAP_CFG 16
#endif
constexpr int WC_CAP = WC_CAP_CFG;         // survivors kept per point
constexpr int WC_NB = 16;                  // histogram bins of d2 / bound
#ifndef WC_MINBLOCKS_CFG
#define WC_MINBLOCKS_CFG 3
#endif

struct __align__(16) WcWarp {
  float4 q[32];                  // moved point; w = squared bound of its search
  u64 list[WC_CAP][32];          // survivors, slot-major (lane = owner reads without bank conflicts)
  uint32_t pref[WC_SEGS + 4];    // first flat candidate index of every compacted segment; pref[nseg] = total
  uint32_t src[WC_SEGS];         // first map_sorted index of the segment
  uint32_t hist[WC_NB / 2][32];  // two 16-bit counters per word
  float scale[32];               // WC_NB / bound
  uint32_t cnt[32];              // survivors appended so far
  int kmax[32];                  // pass 2: last histogram bin that is collected (-1: owner not in pass 2)
  unsigned char own[WC_SEGS];    // owner lane of the segment
};
static_assert(sizeof(u64) * WC_CAP * 32 >= 32 * 8 * sizeof(float), "the staged Jacobian rows reuse the survivor lists");

// W.pref[0 .. nseg) holds the LENGTHS of the warp's compacted segments: turn them into first flat candidate indices
// (exclusive prefix sum, nine consecutive entries per lane) and return the total.
__device__ __forceinline__ uint32_t wc_prefix(WcWarp& W, const int lane, const int nseg) {
  const int e0 = lane * WC_ROWS;
  uint32_t v[WC_ROWS], sum = 0u;
#pragma unroll
  for (int j = 0; j < WC_ROWS; ++j) { v[j] = (e0 + j < nseg) ? W.pref[e0 + j] : 0u; sum += v[j]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t run = incl - sum;
#pragma unroll
  for (int j = 0; j < WC_ROWS; ++j) {
    if (e0 + j < nseg) W.pref[e0 + j] = run;
    run += v[j];
  }
  if (lane == 0) W.pref[nseg] = total;
  __syncwarp();
  return total;
}

// Keep only the segments whose owner takes part in pass 2 (kmax >= 0); returns the new segment count, W.pref holds
// lengths again.  Every lane reads its nine entries before anything is written, and entries only move down.
__device__ __forceinline__ int wc_recompact(WcWarp& W, const int lane, const int nseg) {
  const int e0 = lane * WC_ROWS;
  uint32_t len[WC_ROWS], src[WC_ROWS];
  unsigned char own[WC_ROWS];
  int keep = 0;
#pragma unroll
  for (int j = 0; j < WC_ROWS; ++j) {
    len[j] = 0u; src[j] = 0u; own[j] = 0;
    if (e0 + j < nseg) {
      const int o = W.own[e0 + j];
      if (W.kmax[o] >= 0) { len[j] = W.pref[e0 + j + 1] - W.pref[e0 + j]; src[j] = W.src[e0 + j]; own[j] = (unsigned char)o; ++keep; }
    }
  }
  int incl = keep;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  const int n_new = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  int pos = incl - keep;
#pragma unroll
  for (int j = 0; j < WC_ROWS; ++j) {
    if (len[j] > 0u) { W.pref[pos] = len[j]; W.src[pos] = src[j]; W.own[pos] = own[j]; ++pos; }
  }
  __syncwarp();
  return n_new;
}

// One pass of the warp over its flat candidate list, WC_U loads in flight per lane.  PASS 1: histogram + collect
// everything inside the bound; PASS 2 (list rebuilt from the owners that overflowed): collect the bins <= kmax.
constexpr int WC_U = 4;
template <int PASS>
__device__ __forceinline__ void wc_pass(WcWarp& W, const float4* __restrict__ map_sorted, const int lane, const int nseg,
                                        const uint32_t total) {
  const uint32_t chunk = (total + 31u) >> 5;
  uint32_t c = min((uint32_t)lane * chunk, total);
  const uint32_t c1 = min(c + chunk, total);
  if (c >= c1) return;
  int s = 0;
#pragma unroll
  for (int step = 256; step > 0; step >>= 1)
    if (s + step < nseg && W.pref[s + step] <= c) s += step;
  uint32_t seg_b = W.pref[s], seg_e = W.pref[s + 1], src0 = W.src[s];
  int o = W.own[s];
  while (c < c1) {
    uint32_t a[WC_U];
    int ow[WC_U];
#pragma unroll
    for (int u = 0; u < WC_U; ++u) {
      if (c < c1) {
        if (c >= seg_e) {  // compacted segments are never empty: one step is enough
          ++s;
          seg_b = seg_e; seg_e = W.pref[s + 1]; src0 = W.src[s];
          o = W.own[s];
        }
        a[u] = src0 + (c - seg_b);
        ow[u] = o;
        ++c;
      } else {
        a[u] = a[0];
        ow[u] = -1;
      }
    }
    float4 p[WC_U];
#pragma unroll
    for (int u = 0; u < WC_U; ++u) p[u] = __ldg(map_sorted + a[u]);
#pragma unroll
    for (int u = 0; u < WC_U; ++u) {
      if (ow[u] < 0) continue;
      const float4 q = W.q[ow[u]];
      const float d2 = l2_simple(q, p[u]);
      if (d2 < q.w) {
        const int b = min((int)(d2 * W.scale[ow[u]]), WC_NB - 1);
        if (PASS == 1) atomicAdd(&W.hist[b >> 1][ow[u]], 1u << ((b & 1) * 16));
        if (PASS == 1 || b <= W.kmax[ow[u]]) {
          const uint32_t slot = atomicAdd(&W.cnt[ow[u]], 1u);
          if (slot < (uint32_t)WC_CAP) W.list[slot][ow[u]] = (((u64)__float_as_uint(d2)) << 32) | (u64)__float_as_uint(p[u].w);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(S2M_THREADS, WC_MINBLOCKS_CFG)
s2m_main_wc_kernel(const S2mArgs A) {
  extern __shared__ __align__(16) unsigned char wc_smem[];
  __shared__ float sT[12];
  __shared__ LmTrig sTrig;
  __shared__ double red[S2M_THREADS / 32][S2M_SUMS];
  __shared__ int s_ties, s_wfail[S2M_THREADS / 32], s_iter, s_seeded, s_done0;
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  WcWarp& W = reinterpret_cast<WcWarp*>(wc_smem)[warp];
  cudaGridDependencySynchronize();  // see s2m_main_kernel
  if (tid < 12) sT[tid] = A.T_override ? A.T_override[tid] : A.st->T[tid];
  if (tid == 32) {
    sTrig.srx = A.st->trig[0]; sTrig.crx = A.st->trig[1]; sTrig.sry = A.st->trig[2];
    sTrig.cry = A.st->trig[3]; sTrig.srz = A.st->trig[4]; sTrig.crz = A.st->trig[5];
    s_ties = 0;
    s_seeded = 0;
  }
  if (tid == 64) { s_done0 = A.st->done; s_iter = A.mode == 0 ? A.st->iter : 0; }
  __syncthreads();
  if (A.mode == 0 && s_done0) return;

  const GridParams& g = A.g;
  const int i = blockIdx.x * S2M_THREADS + tid;
  const unsigned lt = (1u << lane) - 1u;
  // ---- A: lane = point ----
  float4 ori = make_float4(0.f, 0.f, 0.f, 0.f), sel = ori;
  int state = 0;  // 0 no point, 1 hopeless (skipped), 2 cooperative search, 3 walks the grid itself
  float bound = g.gate1_d2, bound_seed = FLT_MAX;
  int seeded = 0;
  if (i < A.nq) {
    ori = A.scan[i];
    sel = apply_T(sT, ori);
    state = 2;
    if (s_iter > 0) {
      const int p0 = A.prev_nn[i];
      if (p0 < 0) {
        const float4 hr = A.hopeless[i];
        if (hr.w > 0.f) {
          const float dx = sel.x - hr.x, dy = sel.y - hr.y, dz = sel.z - hr.z;
          if ((dx * dx + dy * dy + dz * dz) * (1.f + HOPELESS_REL) < hr.w * hr.w) state = 1;
        }
      } else {
        const int p1 = A.prev_nn[(size_t)A.nq + i], p2 = A.prev_nn[2 * (size_t)A.nq + i];
        const int p3 = A.prev_nn[3 * (size_t)A.nq + i], p4 = A.prev_nn[4 * (size_t)A.nq + i];
        float D = l2_simple(sel, __ldg(A.map4 + p0));
        D = fmaxf(D, l2_simple(sel, __ldg(A.map4 + p1)));
        D = fmaxf(D, l2_simple(sel, __ldg(A.map4 + p2)));
        D = fmaxf(D, l2_simple(sel, __ldg(A.map4 + p3)));
        D = fmaxf(D, l2_simple(sel, __ldg(A.map4 + p4)));
        const float b = __uint_as_float(__float_as_uint(D) + 1u);  // next float above D: the seeds stay inside
        if (b <= g.gate_d2) seeded = 1;
        // beyond the phase-1 gate the cooperative search uses the gate; if that does not settle the point the seeded
        // ball is searched on its own (see `wide` below)
        bound_seed = b;
        if (b < bound) bound = b;
      }
    }
  }
  uint32_t rs[WC_ROWS], rc[WC_ROWS];
#pragma unroll
  for (int k = 0; k < WC_ROWS; ++k) { rs[k] = 0u; rc[k] = 0u; }
  if (state == 2) {
    const float s2 = 2.0f * g.slack;
    const float reach = sqrtf(bound) * 1.000001f + s2;
    int zmin = (int)floorf((sel.z - reach - g.oz) * g.inv_h), zmax = (int)floorf((sel.z + reach - g.oz) * g.inv_h);
    int ymin = (int)floorf((sel.y - reach - g.oy) * g.inv_h), ymax = (int)floorf((sel.y + reach - g.oy) * g.inv_h);
    zmin = max(zmin, 0); zmax = min(zmax, g.nz - 1);
    ymin = max(ymin, 0); ymax = min(ymax, g.ny - 1);
    if (zmin > zmax || ymin > ymax) {
      // nothing in reach: no segments, fewer than five survivors -> leftover list
    } else if (zmax - zmin > 2 || ymax - ymin > 2) {
      state = 3;
    } else {
      uint32_t tot = 0u;
#pragma unroll
      for (int k = 0; k < WC_ROWS; ++k) {
        const int z = zmin + k / 3, y = ymin + k % 3;
        if (z <= zmax && y <= ymax) {
          const float zlo = g.oz + (float)z * g.h, ylo = g.oy + (float)y * g.h;
          const float gz = fmaxf(fmaxf(zlo - sel.z, sel.z - (zlo + g.h)) - s2, 0.f);
          const float gy = fmaxf(fmaxf(ylo - sel.y, sel.y - (ylo + g.h)) - s2, 0.f);
          const float m2 = (gz * gz + gy * gy) * 0.999999f;
          if (m2 <= bound) {
            const float r = sqrtf(bound - m2) * 1.000001f + s2;
            int xlo = (int)floorf((sel.x - r - g.ox) * g.inv_h);
            int xhi = (int)floorf((sel.x + r - g.ox) * g.inv_h);
            xlo = max(xlo, 0);
            xhi = min(xhi, g.nx - 1);
            if (xlo <= xhi) {
              const uint32_t row = ((uint32_t)z * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nx;
              rs[k] = __ldg(A.cell_start + row + xlo);
              rc[k] = __ldg(A.cell_start + row + xhi + 1);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < WC_ROWS; ++k) { rc[k] -= rs[k]; tot += rc[k]; }
      if (tot > 60000u) state = 3;  // the 16-bit histogram counters must not wrap
    }
  }
  W.q[lane] = make_float4(sel.x, sel.y, sel.z, bound);
  W.scale[lane] = (float)WC_NB / bound;
  W.cnt[lane] = 0u;
  W.kmax[lane] = -1;
#pragma unroll
  for (int k = 0; k < WC_NB / 2; ++k) W.hist[k][lane] = 0u;
  int nseg = 0;
#pragma unroll
  for (int k = 0; k < WC_ROWS; ++k) {
    const bool has = state == 2 && rc[k] > 0u;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (has) {
      const int pos = nseg + __popc(m & lt);
      W.src[pos] = rs[k];
      W.pref[pos] = rc[k];
      W.own[pos] = (unsigned char)lane;
    }
    nseg += __popc(m);
  }
  __syncwarp();
  uint32_t total = wc_prefix(W, lane, nseg);
  // ---- B: lane = candidate ----
  if (total > 0u) {
    wc_pass<1>(W, A.map_sorted, lane, nseg, total);
    __syncwarp();
    uint32_t n = W.cnt[lane];
    bool again = false;
    if (state == 2 && n > (uint32_t)WC_CAP) {
      uint32_t cum = 0u;
      int k5 = -1;
#pragma unroll
      for (int b = 0; b < WC_NB; ++b) {
        cum += (W.hist[b >> 1][lane] >> ((b & 1) * 16)) & 0xffffu;
        if (k5 < 0 && cum >= 5u) { k5 = b; n = cum; }
      }
      if (n <= (uint32_t)WC_CAP) { again = true; W.kmax[lane] = k5; W.cnt[lane] = 0u; }
      else state = 3;  // one bin alone overflows the list
    }
    if (__any_sync(0xffffffffu, again)) {
      __syncwarp();
      const int nseg2 = wc_recompact(W, lane, nseg);
      const uint32_t total2 = wc_prefix(W, lane, nseg2);
      wc_pass<2>(W, A.map_sorted, lane, nseg2, total2);
      __syncwarp();
    }
  }
  // ---- C: lane = point ----
  Top5 t;
  t.init(bound);
  bool need2 = false;
  if (state == 2) {
    const uint32_t n = min(W.cnt[lane], (uint32_t)WC_CAP);
    for (uint32_t j = 0; j < n; ++j) {
      const u64 key = W.list[j][lane];
      t.offer(__uint_as_float((unsigned)(key >> 32)), (int)(unsigned)(key & 0xffffffffull));
    }
    need2 = !(t.d(t.k4) < bound);
  } else if (state == 3) {
    grid_knn5(sel, g, bound, A.map_sorted, A.cell_start, t);
    need2 = !(t.d(t.k4) < bound);
  } else if (state == 1) {
    t.init(g.gate_d2);  // "not found": flag false, no seeds for the next iteration, marker kept
  }
  // A point the phase-1 gate could not settle but whose previous neighbours bound the search inside the full gate
  // (their largest distance to the moved point is a valid radius): the WARP searches that ball for it, lanes taking
  // rows and candidates (warp_knn5) — about 1 us instead of a leftover's full-gate search in the second kernel.
  const bool wide = need2 && bound_seed <= g.gate_d2;
  for (unsigned wm = __ballot_sync(0xffffffffu, wide); wm; wm &= wm - 1u) {
    const int j = __ffs(wm) - 1;
    float4 qj;
    qj.x = __shfl_sync(0xffffffffu, sel.x, j); qj.y = __shfl_sync(0xffffffffu, sel.y, j);
    qj.z = __shfl_sync(0xffffffffu, sel.z, j); qj.w = 0.f;
    const float bj = __shfl_sync(0xffffffffu, bound_seed, j);
    Top5 tj;
    warp_knn5(qj, g, bj, A.map_sorted, A.cell_start, lane, tj, bj);
    if (lane == j) { t = tj; need2 = false; }
  }
  float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float rhs = 0.f;
  bool flag = false, tie = false;
  if (state != 0 && !need2) finish_point(A, i, ori, sel, t, sTrig, row, rhs, flag, tie);
  __syncwarp();
  // ---- the block's sums and leftover segment: as s2m_main_kernel, the staged rows living in the warp's list area ----
  float (*rows)[8] = reinterpret_cast<float (*)[8]>(&W.list[0][0]);
  const unsigned fm = __ballot_sync(0xffffffffu, need2);
  if (lane == 0) s_wfail[warp] = __popc(fm);
#pragma unroll
  for (int k = 0; k < 6; ++k) rows[lane][k] = row[k];
  rows[lane][6] = rhs;
  rows[lane][7] = flag ? 1.f : 0.f;
  if (flag && tie) atomicAdd(&s_ties, 1);
  {
    const int ws = __popc(__ballot_sync(0xffffffffu, seeded != 0));
    if (lane == 0 && ws) atomicAdd(&s_seeded, ws);
  }
  __syncthreads();
  if (need2) {
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_wfail[w];
    A.fail_seg[(size_t)blockIdx.x * S2M_THREADS + base + __popc(fm & lt)] = i;
  }
  {
    const RowAcc ra = row_acc_of(lane);
    double acc = 0.0;
    if (ra.live) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float* rr = rows[r];
        acc += (double)rr[ra.a] * (double)rr[ra.b];
      }
    }
    red[warp][lane] = acc;
  }
  __syncthreads();
  if (tid < S2M_SUMS) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < S2M_THREADS / 32; ++k) sum += red[k][tid];
    if (tid == 28) sum = (double)s_ties;
    if (tid == 29) sum = (double)s_seeded;
    A.partials_main[(size_t)blockIdx.x * S2M_SUMS + tid] = sum;
  }
  if (tid == 0) {
    int nf = 0;
    for (int w = 0; w < S2M_THREADS / 32; ++w) nf += s_wfail[w];
    A.block_nfail[blockIdx.x] = nf;
  }
  if (A.seg_blocks + 1 <= LEFT_OFF_CAP) return;  // the leftover kernel derives the offsets itself
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // ---- last block: exclusive scan of the per-block leftover counts -> fail_off (block order) ----
  __threadfence();
  __shared__ int s_wsum[S2M_THREADS / 32];
  __shared__ int s_carry;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < (int)gridDim.x; b0 += S2M_THREADS) {
    const int b = b0 + tid;
    const int cnt = b < (int)gridDim.x ? __ldcg(A.block_nfail + b) : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int base = s_carry;
    for (int w = 0; w < warp; ++w) base += s_wsum[w];
    if (b < (int)gridDim.x) A.fail_off[b] = base + incl - cnt;
    __syncthreads();
    if (tid == S2M_THREADS - 1) s_carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) {
    A.fail_off[gridDim.x] = s_carry;
    *A.fail_total = s_carry;
    *A.ticket = 0u;
  }
}

constexpr size_t WC_SMEM_BYTES = sizeof(WcWarp) * (S2M_THREADS / 32);

}  // namespace liogpu
