// voxel.cu — sort-based replacement of pcl::VoxelGrid<PointXYZI>::filter as called by liorf at
// mapOptmization.cpp:1536 (key poses), :1582 (local map, leaf surroundingKeyframeMapLeafSize) and
// :1609 (current scan, leaf mappingSurfLeafSize).  Semantics: SURVEY.md Appendix A.1 —
//   min/max (f32, non-finite skipped) -> overflow guard -> int32 voxel index per point -> stable sort
//   by index -> per voxel SEQUENTIAL f32 sums of x,y,z,intensity in ascending input order -> /n.
//
// Kernels (all HBM-bound streaming passes; float4 = one 16-byte coalesced load per point):
//   vox_minmax_kernel   16n B read;  grid = multiple of the SM count, block reduce + ordered-int atomics
//   vox_setup_kernel    1 thread: steps 2-4 of A.1 on device
//   vox_key_kernel      16n read, 4n write
//   radix sort          (sort.cu)  ~3 x 20n per 8-bit pass
//   vox_head_kernel     8n read/write (segment head flags) + exclusive scan
//   vox_segstart_kernel scatter segment starts
//   vox_centroid_kernel one thread per voxel, members gathered through the sorted permutation
#include "common.cuh"

namespace liogpu {

__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ bool finite3(const float4 p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// mm[0..2] = min (ordered uint), mm[3..5] = max, mm[6] = finite count
__global__ void vox_minmax_init_kernel(unsigned* mm) {
  if (threadIdx.x < 3) mm[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) mm[threadIdx.x] = 0u;
  else if (threadIdx.x == 6) mm[6] = 0u;
}

__global__ void __launch_bounds__(256)
vox_minmax_kernel(const float4* __restrict__ pts, int n, unsigned* __restrict__ mm) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (finite3(p)) {
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ float smn[8][3], smx[8][3];
  __shared__ unsigned scnt[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { smn[w][a] = mn[a]; smx[w][a] = mx[a]; }
    scnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    float lo = smn[0][a], hi = smx[0][a];
    for (int k = 1; k < 8; ++k) { lo = fminf(lo, smn[k][a]); hi = fmaxf(hi, smx[k][a]); }
    atomicMin(&mm[a], f2ord(lo));
    atomicMax(&mm[3 + a], f2ord(hi));
  } else if (threadIdx.x == 3) {
    unsigned c = 0;
    for (int k = 0; k < 8; ++k) c += scnt[k];
    atomicAdd(&mm[6], c);
  }
}

// A.1 steps 2-4 from the f32 bounding box and the finite count
__device__ __forceinline__ void vox_setup_compute(const float mn[3], const float mx[3], int n_valid, float leaf,
                                                  VoxelSetup& v) {
  v.n_valid = n_valid;
  const float inv = 1.0f / leaf;
  v.inv_leaf = inv;
  long long d[3];
  for (int a = 0; a < 3; ++a) {
    v.min_p[a] = mn[a];
    v.max_p[a] = mx[a];
    d[a] = (long long)((v.max_p[a] - v.min_p[a]) * inv) + 1;  // A.1 step 3
  }
  v.overflow = (v.n_valid > 0 && d[0] * d[1] * d[2] > 2147483647LL) ? 1 : 0;
  for (int a = 0; a < 3; ++a) {
    v.min_b[a] = (int)floorf(v.min_p[a] * inv);
    const int max_b = (int)floorf(v.max_p[a] * inv);
    v.div_b[a] = max_b - v.min_b[a] + 1;
  }
  v.mul1 = v.div_b[0];
  v.mul2 = v.div_b[0] * v.div_b[1];
  v.n_cells = 0;
  v.key_bits = 0;
  v.n_vox = 0;
  if (!v.overflow && v.n_valid > 0) {
    v.n_cells = (unsigned)v.div_b[0] * (unsigned)v.div_b[1] * (unsigned)v.div_b[2];
    // keys live in [0, n_cells]; n_cells itself marks non-finite points (sorted last, then dropped)
    const unsigned maxkey = v.n_cells;
    int b = 0;
    while (b < 32 && (maxkey >> b) != 0u) ++b;
    v.key_bits = b;
  }
}

__global__ void vox_setup_kernel(const unsigned* __restrict__ mm, float leaf, VoxelSetup* __restrict__ s) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  VoxelSetup v;
  float mn[3], mx[3];
  for (int a = 0; a < 3; ++a) { mn[a] = ord2f(mm[a]); mx[a] = ord2f(mm[3 + a]); }
  vox_setup_compute(mn, mx, (int)mm[6], leaf, v);
  *s = v;
}

__device__ __forceinline__ uint32_t vox_key_of(const float4 p, const VoxelSetup& s) {
  if (!finite3(p)) return s.n_cells;
  const int ix = (int)(floorf(p.x * s.inv_leaf) - (float)s.min_b[0]);  // A.1 step 5
  const int iy = (int)(floorf(p.y * s.inv_leaf) - (float)s.min_b[1]);
  const int iz = (int)(floorf(p.z * s.inv_leaf) - (float)s.min_b[2]);
  return (uint32_t)(ix + iy * s.mul1 + iz * s.mul2);
}

__global__ void __launch_bounds__(256)
vox_key_kernel(const float4* __restrict__ pts, int n, const VoxelSetup* __restrict__ sp, uint32_t* __restrict__ keys) {
  __shared__ VoxelSetup s;
  if (threadIdx.x == 0) s = *sp;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = vox_key_of(pts[i], s);
}

// head[i] = 1 at the first sorted position of every occupied voxel (finite points only), else 0
__global__ void __launch_bounds__(256)
vox_head_kernel(const uint32_t* __restrict__ keys, int n, const VoxelSetup* __restrict__ sp, uint32_t* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int n_valid = sp->overflow ? 0 : sp->n_valid;
  head[i] = (i < n_valid && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}

// seg_start[v] = first sorted position of voxel v; seg_start[n_vox] = n_valid
__global__ void __launch_bounds__(256)
vox_segstart_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, int n,
                    VoxelSetup* __restrict__ sp, uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ n_vox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_valid = sp->overflow ? 0 : sp->n_valid;
  if (i == 0) { seg_start[*n_vox] = (uint32_t)n_valid; sp->n_vox = *n_vox; }
  if (i >= n_valid) return;
  if (i == 0 || keys[i] != keys[i - 1]) seg_start[rank[i]] = (uint32_t)i;
}

// Per-voxel centroid: SEQUENTIAL f32 sums in ascending input index (the sort is stable), true division by
// (float)n — bit-identical to the canonical VoxelGrid (A.1 step 7).  The additions of one voxel cannot be
// reordered, but the LOADS can.
//   vox_centroid_kernel       one thread per voxel; a voxel with <= 8 members is finished here (all member
//                             loads issued before the adds); a longer one is appended to `long_list`.
//   vox_centroid_long_kernel  persistent warps, one long voxel at a time (dynamic: the sizes are heavy
//                             tailed — thousands of members near the sensor when 50 keyframes overlap): the
//                             warp gathers 128 members per step (next step prefetched), stages them in
//                             shared memory, and four lanes — one per component x, y, z, intensity — run the
//                             sequential sums from there (~5 cycles per member instead of one L2 round trip).
constexpr unsigned VOX_SHORT = 8;
constexpr int VOX_LONG_K = 4;                    // 32-member loads per lane and step
constexpr int VOX_LONG_CHUNK = 32 * VOX_LONG_K;  // 128 members per step
constexpr int VOX_LONG_WARPS = 4;

__global__ void __launch_bounds__(256)
vox_centroid_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ perm,
                    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ n_vox_p,
                    float4* __restrict__ out, uint32_t* __restrict__ long_list, uint32_t* __restrict__ long_count) {
  const unsigned nv = *n_vox_p;
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  if ((v & ~31u) >= nv) return;  // whole warp beyond the last voxel
  uint32_t s = 0, e = 0;
  if (v < nv) { s = seg_start[v]; e = seg_start[v + 1]; }
  const uint32_t cnt = e - s;
  const bool is_long = v < nv && cnt > VOX_SHORT;
  if (v < nv && !is_long) {
    float4 p[VOX_SHORT];
#pragma unroll
    for (unsigned k = 0; k < VOX_SHORT; ++k) p[k] = pts[perm[s + (k < cnt ? k : 0)]];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
#pragma unroll
    for (unsigned k = 0; k < VOX_SHORT; ++k)
      if (k < cnt) { sx += p[k].x; sy += p[k].y; sz += p[k].z; si += p[k].w; }
    const float c = (float)cnt;
    out[v] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
  const unsigned lm = __ballot_sync(0xffffffffu, is_long);
  if (lm) {  // one atomic per warp reserves slots for its long voxels (order in the list is irrelevant)
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(long_count, (uint32_t)__popc(lm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is_long) long_list[base + __popc(lm & ((1u << lane) - 1u))] = v;
  }
}

__global__ void __launch_bounds__(32 * VOX_LONG_WARPS)
vox_centroid_long_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ perm,
                         const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ long_list,
                         const uint32_t* __restrict__ long_count, uint32_t* __restrict__ work_counter,
                         float4* __restrict__ out) {
  __shared__ float4 stage[VOX_LONG_WARPS][VOX_LONG_CHUNK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t total = *long_count;
  float4* st = stage[warp];
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const uint32_t v = long_list[item];
    const uint32_t s = seg_start[v], e = seg_start[v + 1];
    float4 cur[VOX_LONG_K], nxt[VOX_LONG_K];
#pragma unroll
    for (int k = 0; k < VOX_LONG_K; ++k) {
      const uint32_t j = s + k * 32 + lane;
      cur[k] = j < e ? pts[perm[j]] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc = 0.f;  // lane c < 4 accumulates component c
    for (uint32_t base = s; base < e; base += VOX_LONG_CHUNK) {
      const uint32_t nb = base + VOX_LONG_CHUNK;
#pragma unroll
      for (int k = 0; k < VOX_LONG_K; ++k) {  // prefetch the next step while this one is consumed
        const uint32_t j = nb + k * 32 + lane;
        nxt[k] = j < e ? pts[perm[j]] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < VOX_LONG_K; ++k) st[k * 32 + lane] = cur[k];
      __syncwarp();
      if (lane < 4) {
        const int m = (int)min((uint32_t)VOX_LONG_CHUNK, e - base);
        const float* f = reinterpret_cast<const float*>(st) + lane;
#pragma unroll 8
        for (int i = 0; i < m; ++i) acc += f[4 * i];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < VOX_LONG_K; ++k) cur[k] = nxt[k];
    }
    const float c = (float)(e - s);
    const float r = acc / c;
    const float rx = __shfl_sync(0xffffffffu, r, 0), ry = __shfl_sync(0xffffffffu, r, 1);
    const float rz = __shfl_sync(0xffffffffu, r, 2), ri = __shfl_sync(0xffffffffu, r, 3);
    if (lane == 0) out[v] = make_float4(rx, ry, rz, ri);
  }
}

// ---- small clouds (the key-pose filter downSizeFilterSurroundingKeyPoses, MO:1535-1536, runs on a few hundred
// poses every scan): the whole filter in ONE block — bounding box, keys, a shared-memory sort of (key, input
// index) pairs (the index makes the order of equal keys the input order), head flags + scan, per-voxel sequential
// sums.  Same arithmetic as the multi-kernel pipeline, one launch instead of twenty.
constexpr int VS_MAX = 2048;
constexpr int VS_THREADS = 1024;
__global__ void __launch_bounds__(VS_THREADS)
vox_small_kernel(const float4* __restrict__ pts, int n, float leaf, VoxelSetup* __restrict__ sp, float4* __restrict__ out) {
  __shared__ unsigned long long comp[VS_MAX];
  __shared__ uint32_t seg[VS_MAX + 1];
  __shared__ float smn[VS_THREADS / 32][3], smx[VS_THREADS / 32][3];
  __shared__ uint32_t swsum[VS_THREADS / 32];
  __shared__ VoxelSetup s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // bounding box of the finite points
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  unsigned cnt = 0;
  for (int i = tid; i < n; i += VS_THREADS) {
    const float4 p = pts[i];
    if (finite3(p)) {
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { smn[w][a] = mn[a]; smx[w][a] = mx[a]; }
    swsum[w] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    unsigned total = 0;
    for (int k = 0; k < VS_THREADS / 32; ++k) {
      total += swsum[k];
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], smn[k][a]); mx[a] = fmaxf(mx[a], smx[k][a]); }
    }
    if (total == 0) {  // same encoding the ordered-uint path decodes for an all-non-finite cloud
      for (int a = 0; a < 3; ++a) { mn[a] = ord2f(0xffffffffu); mx[a] = ord2f(0u); }
    }
    VoxelSetup v;
    vox_setup_compute(mn, mx, (int)total, leaf, v);
    s = v;
  }
  __syncthreads();
  if (s.overflow || s.n_valid == 0) {
    if (tid == 0) *sp = s;
    return;
  }
  int m = 2;
  while (m < n) m <<= 1;  // sort size: next power of two
  for (int i = tid; i < m; i += VS_THREADS)
    comp[i] = i < n ? (((unsigned long long)vox_key_of(pts[i], s) << 32) | (unsigned)i) : ~0ULL;
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = tid; i < m; i += VS_THREADS) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = comp[i], b = comp[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { comp[i] = b; comp[p] = a; }
        }
      }
    }
  }
  __syncthreads();
  // head flags of the occupied voxels (finite points occupy the first n_valid sorted positions) + exclusive scan;
  // thread t owns positions 2t and 2t+1
  const int n_valid = s.n_valid;
  uint32_t h0 = 0, h1 = 0;
  {
    const int i0 = 2 * tid, i1 = 2 * tid + 1;
    if (i0 < n_valid) h0 = (i0 == 0 || (uint32_t)(comp[i0] >> 32) != (uint32_t)(comp[i0 - 1] >> 32)) ? 1u : 0u;
    if (i1 < n_valid) h1 = ((uint32_t)(comp[i1] >> 32) != (uint32_t)(comp[i1 - 1] >> 32)) ? 1u : 0u;
  }
  uint32_t x = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) swsum[w] = x;
  __syncthreads();
  uint32_t wb = 0, total_vox = 0;
  for (int k = 0; k < VS_THREADS / 32; ++k) {
    if (k < w) wb += swsum[k];
    total_vox += swsum[k];
  }
  const uint32_t r0 = wb + x - (h0 + h1);
  if (h0) seg[r0] = 2 * tid;
  if (h1) seg[r0 + h0] = 2 * tid + 1;
  if (tid == 0) seg[total_vox] = (uint32_t)n_valid;
  __syncthreads();
  for (uint32_t v = tid; v < total_vox; v += VS_THREADS) {
    const uint32_t b = seg[v], e = seg[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (uint32_t t = b; t < e; ++t) {  // ascending input index: sequential f32 sums (A.1 step 7)
      const float4 p = pts[(uint32_t)comp[t]];
      sx += p.x; sy += p.y; sz += p.z; si += p.w;
    }
    const float c = (float)(e - b);
    out[v] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
  if (tid == 0) {
    s.n_vox = total_vox;
    *sp = s;
  }
}

// f32 min/max + finite count of a cloud into mm[7] (ordered-uint encoding); shared with grid.cu
cudaError_t launch_minmax(Ctx* c, const float4* pts, int n, unsigned* mm) {
  vox_minmax_init_kernel<<<1, 32, 0, c->stream>>>(mm);
  int grid = div_up(n, 256);
  const int cap = c->sm_count * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  vox_minmax_kernel<<<grid, 256, 0, c->stream>>>(pts, n, mm);
  c->launches += 2;
  return cudaGetLastError();
}

// Device-resident VoxelGrid: in (n float4) -> out (n_out float4).  Everything is enqueued without waiting
// for the host: the key width, the finite count and the overflow flag stay on the device (the radix sort
// runs a fixed four passes, the unneeded ones degrade to a copy); ONE read-back at the end returns the voxel
// count and the guard flag.
int voxel_downsample_dev(Ctx* c, const float4* in, int n, float leaf, DevBuf& out, int* n_out, bool* overflow) {
  *n_out = 0;
  *overflow = false;
  if (n <= 0) return LIOGPU_OK;
  if (!(leaf > 0.f)) { c->err = "voxel leaf must be > 0"; return LIOGPU_E_INVALID; }
  LIOGPU_CUDA_OK(c, c->minmax.reserve(64));
  LIOGPU_CUDA_OK(c, c->vox_setup.reserve(sizeof(VoxelSetup)));
  LIOGPU_CUDA_OK(c, c->keys0.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->keys1.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->vals0.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->vals1.reserve((size_t)n * 4));
  LIOGPU_CUDA_OK(c, c->seg_flag.reserve((size_t)n * 4 + 16));
  LIOGPU_CUDA_OK(c, c->seg_start.reserve((size_t)n * 4 + 16));
  LIOGPU_CUDA_OK(c, c->misc.reserve(256));
  LIOGPU_CUDA_OK(c, out.reserve((size_t)n * sizeof(float4)));
  unsigned* mm = c->minmax.as<unsigned>();
  VoxelSetup* d_setup = c->vox_setup.as<VoxelSetup>();
  VoxelSetup* h_setup = reinterpret_cast<VoxelSetup*>(c->h_pinned);
  if (n <= VS_MAX) {  // small cloud: one block does everything
    vox_small_kernel<<<1, VS_THREADS, 0, c->stream>>>(in, n, leaf, d_setup, out.as<float4>());
    c->launches++;
    LIOGPU_CUDA_OK(c, cudaGetLastError());
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_setup, d_setup, sizeof(VoxelSetup), cudaMemcpyDeviceToHost, c->stream));
    LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (h_setup->overflow) {  // q4
      LIOGPU_CUDA_OK(c, cudaMemcpyAsync(out.p, in, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
      *n_out = n;
      *overflow = true;
      return LIOGPU_OK;
    }
    *n_out = (int)h_setup->n_vox;
    return LIOGPU_OK;
  }
  LIOGPU_CUDA_OK(c, launch_minmax(c, in, n, mm));
  vox_setup_kernel<<<1, 32, 0, c->stream>>>(mm, leaf, d_setup);
  vox_key_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(in, n, d_setup, c->keys0.as<uint32_t>());
  c->launches += 2;
  uint32_t *skeys = nullptr, *sperm = nullptr;
  LIOGPU_CUDA_OK(c, radix_sort_pairs(c, n, -1, &d_setup->key_bits, &skeys, &sperm));
  uint32_t* head = c->seg_flag.as<uint32_t>();
  uint32_t* d_nvox = c->misc.as<uint32_t>();
  vox_head_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(skeys, n, d_setup, head);
  c->launches++;
  LIOGPU_CUDA_OK(c, exclusive_scan_u32(c, head, head, n, d_nvox));
  vox_segstart_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(skeys, head, n, d_setup, c->seg_start.as<uint32_t>(), d_nvox);
  // long-voxel list: reuses the (now dead) unsorted-key ping-pong buffer; two counters in misc
  uint32_t* long_list = (skeys == c->keys0.as<uint32_t>()) ? c->keys1.as<uint32_t>() : c->keys0.as<uint32_t>();
  uint32_t* d_long = c->misc.as<uint32_t>() + 2;  // [0] long count, [1] work counter
  LIOGPU_CUDA_OK(c, cudaMemsetAsync(d_long, 0, 2 * sizeof(uint32_t), c->stream));
  vox_centroid_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(in, sperm, c->seg_start.as<uint32_t>(), d_nvox,
                                                             out.as<float4>(), long_list, d_long);
  vox_centroid_long_kernel<<<c->sm_count * 8, 32 * VOX_LONG_WARPS, 0, c->stream>>>(
      in, sperm, c->seg_start.as<uint32_t>(), long_list, d_long, d_long + 1, out.as<float4>());
  c->launches += 3;
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_setup, d_setup, sizeof(VoxelSetup), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  if (h_setup->overflow) {  // q4: PCL warns and returns the input unchanged
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(out.p, in, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    *n_out = n;
    *overflow = true;
    return LIOGPU_OK;
  }
  *n_out = (int)h_setup->n_vox;
  return LIOGPU_OK;
}

}  // namespace liogpu
