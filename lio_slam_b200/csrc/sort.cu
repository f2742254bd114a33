// sort.cu — hand-written stable LSD radix sort of (u32 key, u32 value) pairs and a u32 exclusive scan.
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
