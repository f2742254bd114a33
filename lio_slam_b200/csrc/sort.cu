// sort.cu — hand-written stable LSD radix sort of (u32 key, u32 value) pairs and a u32 exclusive scan.
// (Round 2: radix_sort_pairs runs the one-sweep kernels further down; the three-launch pass of round 1 is kept as
//  radix_sort_pairs_3launch for A/B — `LIOGPU_SORT=3launch` selects it at run time.)
//
// Used by the VoxelGrid replacement (pcl::VoxelGrid sorts (voxel idx, point idx) pairs; SURVEY A.1
// step 6 — the canonical order is the STABLE one, which is what an LSD radix sort delivers) and by
// order-preserving compaction.  8 bits per pass; ceil(key_bits/8) passes when the host knows the key width,
// otherwise four passes of which the ones beyond the device-side width degrade to a copy (no host round trip).
//
// Per pass, three launches:
//   rs_hist_kernel    per-block digit histogram                     -> counters[digit][block]
//   rs_scan_kernel    one CTA per digit: exclusive scan over blocks -> counters (in place), digit totals
//   rs_scatter_kernel stable multi-split: warp-level match_any ranking in firing order, then scatter
// HBM traffic per pass: read 8n (hist reads keys only: 4n) + read 8n + write 8n bytes.
#include "common.cuh"
#include <stdlib.h>

namespace liogpu {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 2048 keys
constexpr int RS_MAX_BLOCKS = 1024;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, int tiles_per_block,
               uint32_t* __restrict__ counters, int nblocks, const int* __restrict__ d_key_bits) {
  __shared__ uint32_t hist[256];
  if (d_key_bits && shift >= *d_key_bits) return;  // pass not needed for this key range
  hist[threadIdx.x] = 0;
  __syncthreads();
  const long long begin = (long long)blockIdx.x * tiles_per_block * RS_TILE;
  long long end = begin + (long long)tiles_per_block * RS_TILE;
  if (end > n) end = n;
  for (long long i = begin + threadIdx.x; i < end; i += RS_THREADS)
    atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
  __syncthreads();
  counters[(size_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

// One CTA per digit: exclusive scan of that digit's per-block counts (nblocks <= 1024).
__global__ void __launch_bounds__(1024)
rs_scan_kernel(uint32_t* __restrict__ counters, int nblocks, uint32_t* __restrict__ digit_total, int shift,
               const int* __restrict__ d_key_bits) {
  __shared__ uint32_t warp_sum[32];
  if (d_key_bits && shift >= *d_key_bits) return;
  const int d = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  uint32_t v = t < nblocks ? counters[(size_t)d * nblocks + t] : 0u;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sum[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t s = warp_sum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sum[lane] = s;
  }
  __syncthreads();
  const uint32_t incl = x + (w > 0 ? warp_sum[w - 1] : 0u);
  if (t < nblocks) counters[(size_t)d * nblocks + t] = incl - v;
  if (t == 1023) digit_total[d] = incl;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,  // vals_in may be null: iota
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                  int tiles_per_block, const uint32_t* __restrict__ counters, int nblocks,
                  const uint32_t* __restrict__ digit_total, const int* __restrict__ d_key_bits) {
  __shared__ uint32_t warp_hist[RS_WARPS][256];
  if (d_key_bits && shift >= *d_key_bits) {
    // every remaining digit is zero: the pass would be the identity permutation — just move the data
    const long long begin0 = (long long)blockIdx.x * tiles_per_block * RS_TILE;
    long long end0 = begin0 + (long long)tiles_per_block * RS_TILE;
    if (end0 > n) end0 = n;
    for (long long i = begin0 + threadIdx.x; i < end0; i += RS_THREADS) {
      keys_out[i] = keys_in[i];
      vals_out[i] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    return;
  }
  __shared__ uint32_t running[256];
  __shared__ uint32_t wsum[RS_WARPS];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  {  // running[d] = (exclusive scan of digit totals)[d] + this block's offset inside digit d
    const uint32_t v = digit_total[t];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += wsum[k];
    running[t] = base + x - v + counters[(size_t)t * nblocks + blockIdx.x];
  }
  const long long begin = (long long)blockIdx.x * tiles_per_block * RS_TILE;
  long long end = begin + (long long)tiles_per_block * RS_TILE;
  if (end > n) end = n;
  const uint32_t lt_mask = (1u << lane) - 1u;
  for (long long tile = begin; tile < end; tile += RS_TILE) {
#pragma unroll
    for (int k = 0; k < RS_WARPS; ++k) warp_hist[k][t] = 0;
    __syncthreads();
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      const long long idx = tile + (long long)w * (32 * RS_ITEMS) + it * 32 + lane;
      const bool valid = idx < end;
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      key[it] = 0; val[it] = 0; rank[it] = 0;
      if (valid) {
        key[it] = keys_in[idx];
        val[it] = vals_in ? vals_in[idx] : (uint32_t)idx;
        const uint32_t d = (key[it] >> shift) & 255u;
        const unsigned peers = __match_any_sync(act, d);
        const uint32_t pre = warp_hist[w][d];
        __syncwarp(act);
        if ((peers & lt_mask) == 0) warp_hist[w][d] = pre + __popc(peers);
        __syncwarp(act);
        rank[it] = pre + __popc(peers & lt_mask);
      }
    }
    __syncthreads();
    {  // per digit: exclusive prefix over the warps, seeded with the block's running offset
      uint32_t run = running[t];
#pragma unroll
      for (int k = 0; k < RS_WARPS; ++k) {
        const uint32_t c = warp_hist[k][t];
        warp_hist[k][t] = run;
        run += c;
      }
      running[t] = run;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
      const long long idx = tile + (long long)w * (32 * RS_ITEMS) + it * 32 + lane;
      if (idx < end) {
        const uint32_t d = (key[it] >> shift) & 255u;
        const uint32_t pos = warp_hist[w][d] + rank[it];
        keys_out[pos] = key[it];
        vals_out[pos] = val[it];
      }
    }
    __syncthreads();
  }
}

__global__ void iota_kernel(uint32_t* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// Round 2: one-sweep LSD radix sort (chained scan with decoupled look-back).  Same contract as the three-launch
// pass above (stable, 8-bit digits, ping-pong buffers) with ONE launch per pass:
//   os_hist_kernel   one pass over the keys: the digit histograms of ALL passes (shared-memory atomics, then global);
//                    the last block turns each into exclusive digit bases
//   os_pass_kernel   a CTA takes the next tile (atomic ticket, so a tile only ever waits for tiles that already hold
//                    an SM), ranks its 2048 keys (warp-level match_any, firing order), publishes its per-digit counts
//                    and finds its global offsets by looking back at its predecessors' status words
//                    (00 empty | 01 count of this tile | 10 inclusive prefix), stages the tile in digit order in shared
//                    memory and writes every digit's run with consecutive lanes (coalesced), instead of one scattered
//                    4-byte store per key.
// Launches per sort: memset + histogram + passes (6 for 32-bit keys, was 12); bytes per pass: 16n (was 24n).
constexpr uint32_t OS_VALUE_MASK = 0x3fffffffu;
constexpr uint32_t OS_AGGREGATE = 1u << 30, OS_PREFIX = 2u << 30;

__device__ __forceinline__ uint32_t os_ld(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void os_st(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ghist[pass][256] (zeroed); ticket (zeroed).  On return ghist holds exclusive digit bases.
__global__ void __launch_bounds__(RS_THREADS)
os_hist_kernel(const uint32_t* __restrict__ keys, int n, int passes, uint32_t* __restrict__ ghist, uint32_t* __restrict__ ticket) {
  __shared__ uint32_t h[4][256];
  __shared__ bool last;
  const int t = threadIdx.x;
#pragma unroll
  for (int p = 0; p < 4; ++p) h[p][t] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * RS_THREADS + t; i < n; i += (long long)gridDim.x * RS_THREADS) {
    const uint32_t k = keys[i];
    atomicAdd(&h[0][k & 255u], 1u);
    if (passes > 1) atomicAdd(&h[1][(k >> 8) & 255u], 1u);
    if (passes > 2) atomicAdd(&h[2][(k >> 16) & 255u], 1u);
    if (passes > 3) atomicAdd(&h[3][k >> 24], 1u);
  }
  __syncthreads();
  for (int p = 0; p < passes; ++p)
    if (h[p][t]) atomicAdd(&ghist[p * 256 + t], h[p][t]);
  __threadfence();
  __syncthreads();
  if (t == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ uint32_t wsum[RS_WARPS];
  const int lane = t & 31, w = t >> 5;
  for (int p = 0; p < passes; ++p) {  // exclusive scan of the 256 totals of pass p
    const uint32_t v = __ldcg(ghist + p * 256 + t);
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += wsum[k];
    ghist[p * 256 + t] = base + x - v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_THREADS)
os_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,  // vals_in may be null: iota
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
               const uint32_t* __restrict__ gbase, uint32_t* __restrict__ status, uint32_t* __restrict__ tile_ticket,
               const int* __restrict__ d_key_bits) {
  __shared__ uint32_t warp_hist[RS_WARPS][256];
  __shared__ uint32_t s_cnt[256], s_lstart[256], s_gpos[256];
  __shared__ uint32_t s_key[RS_TILE], s_val[RS_TILE];
  __shared__ uint32_t s_wsum[RS_WARPS];
  __shared__ int s_tile;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) s_tile = (int)atomicAdd(tile_ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const long long begin = (long long)tile * RS_TILE;
  long long end = begin + RS_TILE;
  if (end > n) end = n;
  if (d_key_bits && shift >= *d_key_bits) {
    // every remaining digit is zero: the pass would be the identity permutation — just move the data
    for (long long i = begin + t; i < end; i += RS_THREADS) {
      keys_out[i] = keys_in[i];
      vals_out[i] = vals_in ? vals_in[i] : (uint32_t)i;
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < RS_WARPS; ++k) warp_hist[k][t] = 0;
  __syncthreads();
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const long long idx = begin + (long long)w * (32 * RS_ITEMS) + it * 32 + lane;
    const bool valid = idx < end;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    key[it] = 0; val[it] = 0; rank[it] = 0;
    if (valid) {
      key[it] = keys_in[idx];
      val[it] = vals_in ? vals_in[idx] : (uint32_t)idx;
      const uint32_t d = (key[it] >> shift) & 255u;
      // lanes holding the same digit: eight ballots (one per digit bit) instead of match_any, whose cost grows with the
      // number of distinct values in the warp (≈30 for random digits)
      unsigned peers = act;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(act, bit);
        peers &= bit ? m : ~m;
      }
      const uint32_t pre = warp_hist[w][d];
      __syncwarp(act);
      if ((peers & lt_mask) == 0) warp_hist[w][d] = pre + __popc(peers);
      __syncwarp(act);
      rank[it] = pre + __popc(peers & lt_mask);
    }
  }
  __syncthreads();
  // per digit (thread = digit): exclusive prefix over the warps, the tile's count, the chained scan
  uint32_t cnt = 0;
#pragma unroll
  for (int k = 0; k < RS_WARPS; ++k) {
    const uint32_t c = warp_hist[k][t];
    warp_hist[k][t] = cnt;
    cnt += c;
  }
  uint32_t* my_status = status + (size_t)tile * 256 + t;
  os_st(my_status, (tile == 0 ? OS_PREFIX : OS_AGGREGATE) | cnt);
  uint32_t excl = 0;
  for (int p = tile - 1; p >= 0;) {
    const uint32_t v = os_ld(status + (size_t)p * 256 + t);
    const uint32_t flag = v & ~OS_VALUE_MASK;
    if (flag == 0) continue;  // predecessor has not published yet (it holds an SM: the ticket order guarantees progress)
    excl += v & OS_VALUE_MASK;
    if (flag == OS_PREFIX) break;
    --p;
  }
  if (tile > 0) os_st(my_status, OS_PREFIX | (excl + cnt));
  s_cnt[t] = cnt;
  s_gpos[t] = gbase[t] + excl;
  {  // tile-local exclusive scan over the digits
    uint32_t x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[w] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += s_wsum[k];
    s_lstart[t] = base + x - cnt;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const long long idx = begin + (long long)w * (32 * RS_ITEMS) + it * 32 + lane;
    if (idx < end) {
      const uint32_t d = (key[it] >> shift) & 255u;
      const uint32_t lp = s_lstart[d] + warp_hist[w][d] + rank[it];
      s_key[lp] = key[it];
      s_val[lp] = val[it];
    }
  }
  __syncthreads();
  const int tile_n = (int)(end - begin);
  for (int j = t; j < tile_n; j += RS_THREADS) {
    const uint32_t k = s_key[j];
    const uint32_t d = (k >> shift) & 255u;
    const uint32_t pos = s_gpos[d] + ((uint32_t)j - s_lstart[d]);
    keys_out[pos] = k;
    vals_out[pos] = s_val[j];
  }
}

// The round-1 driver (three launches per pass), kept for A/B runs: LIOGPU_SORT=3launch
static cudaError_t radix_sort_pairs_3launch(Ctx* c, int n, int key_bits, const int* d_key_bits, uint32_t** keys_out,
                                            uint32_t** vals_out) {
  uint32_t* k[2] = {c->keys0.as<uint32_t>(), c->keys1.as<uint32_t>()};
  uint32_t* v[2] = {c->vals0.as<uint32_t>(), c->vals1.as<uint32_t>()};
  const int passes = key_bits >= 0 ? (key_bits + 7) / 8 : 4;
  if (key_bits >= 0) d_key_bits = nullptr;
  int cur = 0;
  const int tiles = div_up(n, RS_TILE);
  const int nblocks = tiles < RS_MAX_BLOCKS ? tiles : RS_MAX_BLOCKS;
  const int tiles_per_block = div_up(tiles, nblocks);
  const int nb = div_up(tiles, tiles_per_block);
  cudaError_t e = c->counters.reserve((size_t)(256 * (size_t)nb + 256) * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  uint32_t* counters = c->counters.as<uint32_t>();
  uint32_t* digit_total = counters + (size_t)256 * nb;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    rs_hist_kernel<<<nb, RS_THREADS, 0, c->stream>>>(k[cur], n, shift, tiles_per_block, counters, nb, d_key_bits);
    rs_scan_kernel<<<256, 1024, 0, c->stream>>>(counters, nb, digit_total, shift, d_key_bits);
    rs_scatter_kernel<<<nb, RS_THREADS, 0, c->stream>>>(k[cur], p == 0 ? nullptr : v[cur], k[cur ^ 1], v[cur ^ 1], n,
                                                        shift, tiles_per_block, counters, nb, digit_total, d_key_bits);
    c->launches += 3;
    cur ^= 1;
  }
  *keys_out = k[cur];
  *vals_out = v[cur];
  return cudaGetLastError();
}

// Sort the n pairs whose keys are in c->keys0 (values implicit iota on the first pass).  On return
// *keys_out / *vals_out point at whichever ping-pong buffer holds the result.
// key_bits >= 0: the host knows the key width and runs ceil(key_bits/8) passes.
// key_bits <  0: the width is only known on the device (*d_key_bits); four passes are enqueued and the ones
//                beyond the width degrade to a copy, so no host round trip is needed.
cudaError_t radix_sort_pairs(Ctx* c, int n, int key_bits, const int* d_key_bits, uint32_t** keys_out,
                             uint32_t** vals_out) {
  uint32_t* k[2] = {c->keys0.as<uint32_t>(), c->keys1.as<uint32_t>()};
  uint32_t* v[2] = {c->vals0.as<uint32_t>(), c->vals1.as<uint32_t>()};
  const int passes = key_bits >= 0 ? (key_bits + 7) / 8 : 4;
  if (key_bits >= 0) d_key_bits = nullptr;
  int cur = 0;
  if (n <= 0) { *keys_out = k[0]; *vals_out = v[0]; return cudaSuccess; }
  if (passes == 0) {
    iota_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(v[0], n);
    c->launches++;
    *keys_out = k[0]; *vals_out = v[0];
    return cudaGetLastError();
  }
  static const bool use_3launch = [] { const char* e = getenv("LIOGPU_SORT"); return e && !strcmp(e, "3launch"); }();
  if (use_3launch || n >= (1 << 30)) return radix_sort_pairs_3launch(c, n, key_bits, d_key_bits, keys_out, vals_out);
  const int tiles = div_up(n, RS_TILE);
  // scratch: [4][256] digit bases | 8 tickets | [passes][tiles][256] status words, zeroed in one memset
  const size_t words = 4 * 256 + 8 + (size_t)passes * (size_t)tiles * 256;
  cudaError_t e = c->counters.reserve(words * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  uint32_t* ghist = c->counters.as<uint32_t>();
  uint32_t* tickets = ghist + 4 * 256;
  uint32_t* status = tickets + 8;
  e = cudaMemsetAsync(ghist, 0, words * sizeof(uint32_t), c->stream);
  if (e != cudaSuccess) return e;
  int hist_blocks = tiles < c->sm_count * 8 ? tiles : c->sm_count * 8;
  os_hist_kernel<<<hist_blocks, RS_THREADS, 0, c->stream>>>(k[0], n, passes, ghist, tickets);
  c->launches++;
  for (int p = 0; p < passes; ++p) {
    os_pass_kernel<<<tiles, RS_THREADS, 0, c->stream>>>(k[cur], p == 0 ? nullptr : v[cur], k[cur ^ 1], v[cur ^ 1], n, 8 * p,
                                                       ghist + p * 256, status + (size_t)p * tiles * 256, tickets + 1 + p,
                                                       d_key_bits);
    c->launches++;
    cur ^= 1;
  }
  *keys_out = k[cur];
  *vals_out = v[cur];
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan (u32) — block scan of 2048-element chunks, recursive scan of the chunk totals, add.
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__global__ void __launch_bounds__(SC_THREADS)
scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n, uint32_t* __restrict__ tile_total) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const long long base = (long long)blockIdx.x * SC_TILE + (long long)t * SC_ITEMS;
  uint32_t v[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  uint32_t wb = 0;
  for (int k = 0; k < w; ++k) wb += wsum[k];
  uint32_t run = wb + x - s;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (t == SC_THREADS - 1 && tile_total) tile_total[blockIdx.x] = run;
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, int n, const uint32_t* __restrict__ tile_base) {
  const long long i = (long long)blockIdx.x * SC_TILE + threadIdx.x;
  const uint32_t b = tile_base[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    const long long j = i + (long long)k * SC_THREADS;
    if (j < n) out[j] += b;
  }
}

static cudaError_t scan_rec(Ctx* c, const uint32_t* in, uint32_t* out, int n, uint32_t* tmp, uint32_t* d_total) {
  const int tiles = div_up(n, SC_TILE);
  if (tiles <= 1) {
    scan_tile_kernel<<<1, SC_THREADS, 0, c->stream>>>(in, out, n, d_total);
    c->launches++;
    return cudaGetLastError();
  }
  // tmp layout: [tiles totals][tiles scanned totals][rest for deeper levels]
  uint32_t* totals = tmp;
  uint32_t* scanned = tmp + tiles;
  scan_tile_kernel<<<tiles, SC_THREADS, 0, c->stream>>>(in, out, n, totals);
  c->launches++;
  cudaError_t e = scan_rec(c, totals, scanned, tiles, tmp + 2 * (size_t)tiles, d_total);
  if (e != cudaSuccess) return e;
  scan_add_kernel<<<tiles, SC_THREADS, 0, c->stream>>>(out, n, scanned);
  c->launches++;
  return cudaGetLastError();
}

// out[i] = sum_{j<i} in[j]; *d_total (device, optional) = sum of all.  in == out is allowed.
cudaError_t exclusive_scan_u32(Ctx* c, const uint32_t* in, uint32_t* out, int n, uint32_t* d_total) {
  if (n <= 0) {
    if (d_total) return cudaMemsetAsync(d_total, 0, sizeof(uint32_t), c->stream);
    return cudaSuccess;
  }
  const size_t tiles = (size_t)div_up(n, SC_TILE);
  cudaError_t e = c->scan_tmp.reserve((tiles * 3 + 64) * sizeof(uint32_t));
  if (e != cudaSuccess) return e;
  return scan_rec(c, in, out, n, c->scan_tmp.as<uint32_t>(), d_total);
}

}  // namespace liogpu
