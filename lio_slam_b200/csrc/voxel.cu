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

__global__ void vox_setup_kernel(const unsigned* __restrict__ mm, float leaf, VoxelSetup* __restrict__ s) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  VoxelSetup v;
  v.n_valid = (int)mm[6];
  const float inv = 1.0f / leaf;
  v.inv_leaf = inv;
  long long d[3];
  for (int a = 0; a < 3; ++a) {
    v.min_p[a] = ord2f(mm[a]);
    v.max_p[a] = ord2f(mm[3 + a]);
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
  if (!v.overflow && v.n_valid > 0) {
    v.n_cells = (unsigned)v.div_b[0] * (unsigned)v.div_b[1] * (unsigned)v.div_b[2];
    // keys live in [0, n_cells]; n_cells itself marks non-finite points (sorted last, then dropped)
    const unsigned maxkey = v.n_cells;
    int b = 0;
    while (b < 32 && (maxkey >> b) != 0u) ++b;
    v.key_bits = b;
  }
  *s = v;
}

__global__ void __launch_bounds__(256)
vox_key_kernel(const float4* __restrict__ pts, int n, const VoxelSetup* __restrict__ sp, uint32_t* __restrict__ keys) {
  __shared__ VoxelSetup s;
  if (threadIdx.x == 0) s = *sp;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  uint32_t key = s.n_cells;
  if (finite3(p)) {
    const int ix = (int)(floorf(p.x * s.inv_leaf) - (float)s.min_b[0]);  // A.1 step 5
    const int iy = (int)(floorf(p.y * s.inv_leaf) - (float)s.min_b[1]);
    const int iz = (int)(floorf(p.z * s.inv_leaf) - (float)s.min_b[2]);
    key = (uint32_t)(ix + iy * s.mul1 + iz * s.mul2);
  }
  keys[i] = key;
}

__global__ void __launch_bounds__(256)
vox_head_kernel(const uint32_t* __restrict__ keys, int n_valid, uint32_t* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// seg_start[v] = first sorted position of voxel v; seg_start[n_vox] = n_valid
__global__ void __launch_bounds__(256)
vox_segstart_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, int n_valid,
                    uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ n_vox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) seg_start[*n_vox] = (uint32_t)n_valid;
  if (i >= n_valid) return;
  if (i == 0 || keys[i] != keys[i - 1]) seg_start[rank[i]] = (uint32_t)i;
}

// One thread per voxel: sequential f32 sums in ascending input index (the sort is stable), true
// division by (float)n — bit-identical to the canonical VoxelGrid (A.1 step 7).
__global__ void __launch_bounds__(256)
vox_centroid_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ perm,
                    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ n_vox_p,
                    float4* __restrict__ out) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= *n_vox_p) return;
  const uint32_t s = seg_start[v], e = seg_start[v + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (uint32_t j = s; j < e; ++j) {
    const float4 p = pts[perm[j]];
    sx += p.x; sy += p.y; sz += p.z; si += p.w;
  }
  const float cnt = (float)(e - s);
  out[v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
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

// Device-resident VoxelGrid: in (n float4) -> out (n_out float4).  One small D2H sync for the setup
// block (the host must size the sort) and one for the voxel count.
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
  unsigned* mm = c->minmax.as<unsigned>();
  VoxelSetup* d_setup = c->vox_setup.as<VoxelSetup>();
  LIOGPU_CUDA_OK(c, launch_minmax(c, in, n, mm));
  vox_setup_kernel<<<1, 32, 0, c->stream>>>(mm, leaf, d_setup);
  c->launches += 1;
  VoxelSetup* h_setup = reinterpret_cast<VoxelSetup*>(c->h_pinned);
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_setup, d_setup, sizeof(VoxelSetup), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  const VoxelSetup hs = *h_setup;
  if (hs.overflow) {  // q4: PCL warns and returns the input unchanged
    LIOGPU_CUDA_OK(c, out.reserve((size_t)n * sizeof(float4)));
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(out.p, in, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
    *n_out = n;
    *overflow = true;
    return LIOGPU_OK;
  }
  if (hs.n_valid <= 0) return LIOGPU_OK;
  vox_key_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(in, n, d_setup, c->keys0.as<uint32_t>());
  c->launches++;
  uint32_t *skeys = nullptr, *sperm = nullptr;
  LIOGPU_CUDA_OK(c, radix_sort_pairs(c, n, hs.key_bits, &skeys, &sperm));
  uint32_t* head = c->seg_flag.as<uint32_t>();
  uint32_t* d_nvox = c->misc.as<uint32_t>();
  vox_head_kernel<<<div_up(hs.n_valid, 256), 256, 0, c->stream>>>(skeys, hs.n_valid, head);
  c->launches++;
  LIOGPU_CUDA_OK(c, exclusive_scan_u32(c, head, head, hs.n_valid, d_nvox));
  vox_segstart_kernel<<<div_up(hs.n_valid, 256), 256, 0, c->stream>>>(skeys, head, hs.n_valid,
                                                                       c->seg_start.as<uint32_t>(), d_nvox);
  c->launches++;
  // upper bound of the voxel count is n_valid: launch for that and let surplus threads exit
  LIOGPU_CUDA_OK(c, out.reserve((size_t)hs.n_valid * sizeof(float4)));
  vox_centroid_kernel<<<div_up(hs.n_valid, 256), 256, 0, c->stream>>>(in, sperm, c->seg_start.as<uint32_t>(), d_nvox,
                                                                      out.as<float4>());
  c->launches++;
  uint32_t* h_nvox = reinterpret_cast<uint32_t*>((char*)c->h_pinned + 1024);
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_nvox, d_nvox, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  *n_out = (int)*h_nvox;
  return LIOGPU_OK;
}

}  // namespace liogpu
