// s2m_split.cuh — the main kernel of the two-kernel LM loop cut in two (LIOGPU_MAIN=split), included by s2m.cu:
//
//   s2m_search_kernel  the search part of s2m_main_kernel only (transform, seeded / phase-1 grid walk, leftover list):
//                      no plane fit, no Jacobian, no FP64 sums -> far fewer live registers, so more warps are resident to
//                      hide the walk's dependent look-ups (ncu on s2m_main_kernel: 27 % of the stall samples on the long
//                      scoreboard, 17 % on fixed-latency dependencies, 40 % occupancy at 64 registers).  It leaves the
//                      five neighbour indices in prev_nn — which the next iteration's seeded search needs anyway — with
//                      the tie bit of the point in bit 30 of the fifth index, and -2 for a point that went to the
//                      leftover list.
//   s2m_fit_kernel     one thread per point, fully convergent: neighbour gather, 5x3 plane fit, Jacobian row, FP64 block
//                      sums into the same partial rows s2m_main_kernel writes.
//
// STATUS: A/B variant (LIOGPU_MAIN=split), parity-green (tests/test_gpu_variants.py), MEASURED NOT FASTER: search + fit
// 70-79 us per iteration on config 3 (s2m_main_kernel: 72 before, 68 after the collecting walk) — the one-wave residency
// did not shorten the search, and the second launch + re-gather of the neighbours cost what it saved.
//
// The leftover kernel follows unchanged.  Results are bit-identical to s2m_main_kernel (same per-point arithmetic, same
// assignment of points to partial rows).
#pragma once

namespace liogpu {

constexpr int NN_TIE_BIT = 1 << 30;     // fifth neighbour index: the point's five-nearest set has an equidistant tie
constexpr int NN_SEEDED_BIT = 1 << 29;  // fifth neighbour index: the search started from the previous neighbours (statistic)
constexpr int NN_INDEX_MASK = NN_SEEDED_BIT - 1;
constexpr int NN_PENDING = -2;  // the point is on the leftover list: s2m_left_kernel will write its neighbours

// ONE WAVE: with 64-thread blocks at <= 40 registers, 25 blocks (50 warps) fit an SM, so the 3,600 blocks of a 230,400-point
// sweep are all resident at once on 148 SMs.  A point's search is a chain of dependent look-ups, so a launch costs
// (waves) x (chain latency) almost independently of occupancy: the fused main kernel needs 1.5 waves at 32 warps per SM,
// i.e. two chain latencies; the search alone, in one wave, needs one.
#ifndef SEARCH_THREADS_CFG
#define SEARCH_THREADS_CFG 64
#endif
#ifndef SEARCH_MINBLOCKS_CFG
#define SEARCH_MINBLOCKS_CFG 24
#endif
constexpr int SEARCH_THREADS = SEARCH_THREADS_CFG;

__global__ void __launch_bounds__(SEARCH_THREADS, SEARCH_MINBLOCKS_CFG)
s2m_search_kernel(const S2mArgs A) {
  __shared__ float sT[12];
  __shared__ int s_wfail[SEARCH_THREADS / 32], s_iter, s_done0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cudaGridDependencySynchronize();
  if (tid < 12) sT[tid] = A.T_override ? A.T_override[tid] : A.st->T[tid];
  if (tid == 32) { s_done0 = A.st->done; s_iter = A.mode == 0 ? A.st->iter : 0; }
  __syncthreads();
  if (A.mode == 0 && s_done0) return;
  const int i = blockIdx.x * SEARCH_THREADS + tid;
  bool need2 = false;
  if (i < A.nq) {
    const float4 sel = apply_T(sT, A.scan[i]);
    Top5 t;
    float gate_use = A.g.gate1_d2;
    bool can_search = A.g.gate1_d2 < A.g.gate_d2, is_seeded = false, skip = false;
    if (s_iter > 0) {
      const int p0 = A.prev_nn[i];
      if (p0 < 0) {
        const float4 hr = A.hopeless[i];
        if (hr.w > 0.f) {
          const float dx = sel.x - hr.x, dy = sel.y - hr.y, dz = sel.z - hr.z;
          skip = (dx * dx + dy * dy + dz * dz) * (1.f + HOPELESS_REL) < hr.w * hr.w;
        }
      } else {
        const int p1 = A.prev_nn[(size_t)A.nq + i], p2 = A.prev_nn[2 * (size_t)A.nq + i];
        const int p3 = A.prev_nn[3 * (size_t)A.nq + i], p4 = A.prev_nn[4 * (size_t)A.nq + i] & NN_INDEX_MASK;
        const float d0 = l2_simple(sel, __ldg(A.map4 + p0)), d1 = l2_simple(sel, __ldg(A.map4 + p1));
        const float d2 = l2_simple(sel, __ldg(A.map4 + p2)), d3 = l2_simple(sel, __ldg(A.map4 + p3));
        const float d4 = l2_simple(sel, __ldg(A.map4 + p4));
        const float D = fmaxf(fmaxf(fmaxf(d0, d1), fmaxf(d2, d3)), d4);
        const float bound = __uint_as_float(__float_as_uint(D) + 1u);
        if (bound <= A.g.gate_d2) {
          gate_use = bound; can_search = true; is_seeded = true;
#ifndef SPLIT_NO_SEED_INIT
          // the five seeds enter the list up front (all lanes together): in a late iteration they ARE the answer and the
          // walk then inserts nothing — the insertion network otherwise runs ~30 times per warp at 5 of 32 lanes
          t.init(bound);
          t.offer(d0, p0); t.offer(d1, p1); t.offer(d2, p2); t.offer(d3, p3); t.offer(d4, p4);
#endif
        }
      }
    }
    if (skip) {
      t.init(A.g.gate_d2);
    } else if (can_search) {
#ifndef SPLIT_NO_SEED_INIT
      if (is_seeded) {
#ifdef SPLIT_BOX9  // measured slower on the dense map (27 more live registers: spills; 100 vs 70 us per launch)
        // a seeded bound is tight from the start, so nothing is lost by fixing every row's x range up front: all
        // cell-table look-ups of the (at most 3 x 3) rows are in flight together instead of one round trip per row
        if (!grid_knn5_box9<true>(sel, A.g, gate_use, A.map_sorted, A.cell_start, t))
#endif
          grid_knn5<true>(sel, A.g, gate_use, A.map_sorted, A.cell_start, t);
      } else
#endif
        grid_knn5(sel, A.g, gate_use, A.map_sorted, A.cell_start, t);
      need2 = !is_seeded && !(t.d(t.k4) < A.g.gate1_d2);
    } else {
      need2 = true;
    }
    const bool found = !need2 && t.d(t.k4) < A.g.gate_d2;
    A.prev_nn[i] = need2 ? NN_PENDING : (found ? t.i(t.k0) : -1);
    if (found) {
      A.prev_nn[(size_t)A.nq + i] = t.i(t.k1);
      A.prev_nn[2 * (size_t)A.nq + i] = t.i(t.k2);
      A.prev_nn[3 * (size_t)A.nq + i] = t.i(t.k3);
      A.prev_nn[4 * (size_t)A.nq + i] = t.i(t.k4) | (t.tie() ? NN_TIE_BIT : 0) | (is_seeded ? NN_SEEDED_BIT : 0);
    }
    if (A.mode == 1 && !need2) {
      if (A.dbg.nn_idx) {
        int* o = A.dbg.nn_idx + (size_t)i * 5;
        o[0] = found ? t.i(t.k0) : -1; o[1] = found ? t.i(t.k1) : -1; o[2] = found ? t.i(t.k2) : -1;
        o[3] = found ? t.i(t.k3) : -1; o[4] = found ? t.i(t.k4) : -1;
      }
      if (A.dbg.nn_d2) {
        float* o = A.dbg.nn_d2 + (size_t)i * 5;
        o[0] = t.d(t.k0); o[1] = t.d(t.k1); o[2] = t.d(t.k2); o[3] = t.d(t.k3); o[4] = t.d(t.k4);
      }
    }
  }
  const unsigned fm = __ballot_sync(0xffffffffu, need2);
  if (lane == 0) s_wfail[warp] = __popc(fm);
  __syncthreads();
  if (need2) {
    int base = 0;
    for (int w = 0; w < warp; ++w) base += s_wfail[w];
    A.fail_seg[(size_t)blockIdx.x * SEARCH_THREADS + base + __popc(fm & ((1u << lane) - 1u))] = i;
  }
  if (tid == 0) {
    int nf = 0;
    for (int w = 0; w < SEARCH_THREADS / 32; ++w) nf += s_wfail[w];
    A.block_nfail[blockIdx.x] = nf;
  }
}

__global__ void __launch_bounds__(S2M_THREADS, 4)
s2m_fit_kernel(const S2mArgs A) {
  __shared__ float sT[12];
  __shared__ LmTrig sTrig;
  __shared__ float rows[S2M_THREADS][8];
  __shared__ double red[S2M_THREADS / 32][S2M_SUMS];
  __shared__ int s_ties, s_seeded, s_done0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cudaGridDependencySynchronize();
  if (tid < 12) sT[tid] = A.T_override ? A.T_override[tid] : A.st->T[tid];
  if (tid == 32) {
    sTrig.srx = A.st->trig[0]; sTrig.crx = A.st->trig[1]; sTrig.sry = A.st->trig[2];
    sTrig.cry = A.st->trig[3]; sTrig.srz = A.st->trig[4]; sTrig.crz = A.st->trig[5];
    s_ties = 0;
    s_seeded = 0;
  }
  if (tid == 64) s_done0 = A.st->done;
  __syncthreads();
  if (A.mode == 0 && s_done0) return;
  const int i = blockIdx.x * S2M_THREADS + tid;
  float row[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float rhs = 0.f;
  bool flag = false, tie = false, seeded = false;
  if (i < A.nq) {
    const int p0 = A.prev_nn[i];
    if (p0 != NN_PENDING) {
      float4 coeff = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p0 >= 0) {
        const int p1 = A.prev_nn[(size_t)A.nq + i], p2 = A.prev_nn[2 * (size_t)A.nq + i];
        const int p3 = A.prev_nn[3 * (size_t)A.nq + i], p4t = A.prev_nn[4 * (size_t)A.nq + i];
        const float4 ori = A.scan[i];
        const float4 sel = apply_T(sT, ori);
        float4 nbr[5];
        nbr[0] = __ldg(A.map4 + p0); nbr[1] = __ldg(A.map4 + p1); nbr[2] = __ldg(A.map4 + p2);
        nbr[3] = __ldg(A.map4 + p3); nbr[4] = __ldg(A.map4 + (p4t & NN_INDEX_MASK));
        flag = plane_residual(ori, sel, nbr, coeff);
        tie = (p4t & NN_TIE_BIT) != 0;
        if (!flag) coeff = make_float4(0.f, 0.f, 0.f, 0.f);
        if (flag) jacobian_row(sTrig, ori, coeff, row, rhs);
        seeded = (p4t & NN_SEEDED_BIT) != 0;
      }
      if (A.mode == 1) {
        if (A.dbg.coeff) A.dbg.coeff[i] = coeff;
        if (A.dbg.flag) A.dbg.flag[i] = flag ? 1 : 0;
        if (A.dbg.tie) A.dbg.tie[i] = tie ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) rows[tid][k] = row[k];
  rows[tid][6] = rhs;
  rows[tid][7] = flag ? 1.f : 0.f;
  if (flag && tie) atomicAdd(&s_ties, 1);
  {
    const int ws = __popc(__ballot_sync(0xffffffffu, seeded));
    if (lane == 0 && ws) atomicAdd(&s_seeded, ws);
  }
  __syncthreads();
  {
    const RowAcc ra = row_acc_of(lane);
    double acc = 0.0;
    if (ra.live) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float* rr = rows[warp * 32 + r];
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
}

}  // namespace liogpu
