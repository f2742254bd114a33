// layout.cu — record <-> packed float4 conversion and the rigid transform of a cloud.
//
// On device every cloud is a dense float4 array (x, y, z, intensity): one 16-byte vector load per
// point, fully coalesced.  The reference's 32-byte pcl::PointXYZI records (utility.h:65) are unpacked
// once at the boundary.
#include "common.cuh"
#include "pose_math.cuh"

namespace liogpu {

// 32-byte records: each thread reads two float4 halves (x,y,z,pad | intensity,...) — coalesced 16 B loads.
__global__ void unpack_kernel(const unsigned char* __restrict__ raw, int n, int stride, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* r = raw + (size_t)i * stride;
  if (stride == 16) {
    out[i] = *reinterpret_cast<const float4*>(r);
  } else if ((stride & 15) == 0) {
    const float4 a = *reinterpret_cast<const float4*>(r);
    const float4 b = *reinterpret_cast<const float4*>(r + 16);
    out[i] = make_float4(a.x, a.y, a.z, b.x);
  } else {
    const float* f = reinterpret_cast<const float*>(r);
    out[i] = make_float4(f[0], f[1], f[2], f[4]);
  }
}

__global__ void pack_kernel(const float4* __restrict__ in, int n, unsigned char* __restrict__ raw, int stride) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  unsigned char* r = raw + (size_t)i * stride;
  if (stride == 16) {
    *reinterpret_cast<float4*>(r) = p;
  } else if ((stride & 15) == 0) {
    *reinterpret_cast<float4*>(r) = make_float4(p.x, p.y, p.z, 1.0f);  // PCL keeps data[3] = 1
    *reinterpret_cast<float4*>(r + 16) = make_float4(p.w, 0.f, 0.f, 0.f);
    for (int o = 32; o + 16 <= stride; o += 16) *reinterpret_cast<float4*>(r + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    float* f = reinterpret_cast<float*>(r);
    f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = 1.0f; f[4] = p.w;
    for (int o = 5; o < stride / 4; ++o) f[o] = 0.f;
  }
}

// mapOptimization::transformPointCloud (mapOptmization.cpp:849-868): out = R*p + t, f32, no FMA,
// products summed left to right exactly as written at :862-864.
__global__ void transform_kernel(const float4* __restrict__ in, int n, const float* __restrict__ pose6,
                                 float4* __restrict__ out) {
  __shared__ float T[12];
  if (threadIdx.x == 0) pose_to_T(pose6, T);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = apply_T(T, in[i]);
}

// extractCloud (mapOptmization.cpp:1556-1588): all keyframes of a local map in ONE launch instead of one
// transformPointCloud per keyframe.  pose_table_kernel builds the k 3x4 transforms; transform_multi_kernel
// finds the keyframe of every output point by binary search in the offset table (k <= 1365, L1-resident).
__global__ void pose_table_kernel(const float* __restrict__ poses6, int k, float* __restrict__ T12) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= k) return;
  float T[12];
  pose_to_T(poses6 + 6 * f, T);
#pragma unroll
  for (int q = 0; q < 12; ++q) T12[12 * f + q] = T[q];
}
__global__ void __launch_bounds__(256)
transform_multi_kernel(const float4* const* __restrict__ srcs, const int* __restrict__ offs, int k,
                       const float* __restrict__ T12, long long total, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = k;  // offs[lo] <= i < offs[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(offs + mid) <= i) lo = mid; else hi = mid;
  }
  const float4 p = srcs[lo][i - __ldg(offs + lo)];
  float T[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) T[q] = __ldg(T12 + 12 * lo + q);
  out[i] = apply_T(T, p);
}
cudaError_t launch_pose_table(Ctx* c, const float* d_poses6, int k, float* d_T12) {
  if (k <= 0) return cudaSuccess;
  pose_table_kernel<<<div_up(k, 128), 128, 0, c->stream>>>(d_poses6, k, d_T12);
  c->launches++;
  return cudaGetLastError();
}
cudaError_t launch_transform_multi(Ctx* c, const float4* const* d_srcs, const int* d_offs, int k, const float* d_poses6,
                                   float* d_T12, long long total, float4* out) {
  if (k <= 0 || total <= 0) return cudaSuccess;
  cudaError_t e = launch_pose_table(c, d_poses6, k, d_T12);
  if (e != cudaSuccess) return e;
  transform_multi_kernel<<<div_up(total, 256), 256, 0, c->stream>>>(d_srcs, d_offs, k, d_T12, total, out);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_unpack(Ctx* c, const void* d_raw, int n, int stride, float4* out) {
  if (n <= 0) return cudaSuccess;
  unpack_kernel<<<div_up(n, 256), 256, 0, c->stream>>>((const unsigned char*)d_raw, n, stride, out);
  c->launches++;
  return cudaGetLastError();
}
cudaError_t launch_pack(Ctx* c, const float4* in, int n, void* d_raw, int stride) {
  if (n <= 0) return cudaSuccess;
  pack_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(in, n, (unsigned char*)d_raw, stride);
  c->launches++;
  return cudaGetLastError();
}
cudaError_t launch_transform(Ctx* c, const float4* in, int n, const float* d_pose6, float4* out) {
  if (n <= 0) return cudaSuccess;
  transform_kernel<<<div_up(n, 256), 256, 0, c->stream>>>(in, n, d_pose6, out);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace liogpu
