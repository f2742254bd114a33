// deskew.cu — ImageProjection::projectPointCloud + deskewPoint + findRotation
// (imageProjection.cpp:577-615, 545-575, 502-527) as three streaming passes:
//   dsk_flag_kernel   32n B read: crop box / range / intensity / ring / decimation tests (:596-609),
//                     survivor flag per point, atomicMin of the first survivor (quirk q6)
//   exclusive scan    order-preserving output slot per survivor (sort.cu)
//   dsk_apply_kernel  per survivor: IMU rotation by linear interpolation in f64 (findRotation),
//                     R = Rz*Ry*Rx in f32, transBt = transStartInverse * transFinal, p' = transBt*p
// The IMU table (imuTime, imuRotX/Y/Z; <= 2000 rows, IP:62) is read through the read-only path; every
// thread of a warp touches the same few rows.  Positional deskew is disabled in the reference
// (findPosition returns zeros, :529-543), so translations are exactly 0.
#include "common.cuh"
#include "pose_math.cuh"

namespace liogpu {

struct DeskewConst {
  int n_scan, downsample_rate, point_filter_num;
  float min_front, min_back, min_left, min_right, max_range, max_intensity;
  double t_scan;
  int n_imu;
  int enabled;
};

struct RawPoint {
  float x, y, z, intensity, time;
  int ring;
};

__device__ __forceinline__ RawPoint load_raw(const unsigned char* __restrict__ raw, int i, int stride) {
  const unsigned char* r = raw + (size_t)i * stride;
  RawPoint p;
  if ((stride & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(r));
    const float4 b = __ldg(reinterpret_cast<const float4*>(r + 16));
    p.x = a.x; p.y = a.y; p.z = a.z;
    p.intensity = b.x;
    p.ring = (int)(__float_as_uint(b.y) & 0xffffu);  // uint16 ring at byte 20 (little endian)
    p.time = b.z;
  } else {
    const float* f = reinterpret_cast<const float*>(r);
    p.x = f[0]; p.y = f[1]; p.z = f[2]; p.intensity = f[4];
    p.ring = (int)(__float_as_uint(f[5]) & 0xffffu);
    p.time = f[6];
  }
  return p;
}

__global__ void __launch_bounds__(256)
dsk_flag_kernel(const unsigned char* __restrict__ raw, int n, int stride, const DeskewConst k,
                uint32_t* __restrict__ flags, int* __restrict__ first_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RawPoint p = load_raw(raw, i, stride);
  const float range = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);  // common_lib pointDistance
  bool keep = !((p.y < k.min_front && -k.min_back < p.y && p.x < k.min_left && -k.min_right < p.x) ||
                range > k.max_range || p.intensity > k.max_intensity);
  keep = keep && (p.ring >= 0 && p.ring < k.n_scan);
  keep = keep && (p.ring % k.downsample_rate == 0);
  keep = keep && (i % k.point_filter_num == 0);  // raw index, before cropping (quirk q5)
  flags[i] = keep ? 1u : 0u;
  if (keep) atomicMin(first_idx, i);
}

// findRotation (:502-527): linear scan for the first row whose time exceeds pointTime, lerp in f64.
__device__ __forceinline__ void find_rotation(double t, const double* __restrict__ tab, int n_imu, float& rx,
                                              float& ry, float& rz) {
  const double* imu_t = tab;
  const double* ax = tab + n_imu;
  const double* ay = tab + 2 * n_imu;
  const double* az = tab + 3 * n_imu;
  const int ptr_cur = n_imu - 1;
  int f = 0;
  while (f < ptr_cur) {
    if (t < __ldg(imu_t + f)) break;
    ++f;
  }
  const double tf = __ldg(imu_t + f);
  if (t > tf || f == 0) {
    rx = (float)__ldg(ax + f); ry = (float)__ldg(ay + f); rz = (float)__ldg(az + f);
  } else {
    const int b = f - 1;
    const double tb = __ldg(imu_t + b);
    const double rf = (t - tb) / (tf - tb);
    const double rb = (tf - t) / (tf - tb);
    rx = (float)(__ldg(ax + f) * rf + __ldg(ax + b) * rb);
    ry = (float)(__ldg(ay + f) * rf + __ldg(ay + b) * rb);
    rz = (float)(__ldg(az + f) * rf + __ldg(az + b) * rb);
  }
}

__device__ __forceinline__ void rot_from_rpy(float rx, float ry, float rz, float R[9]) {
  const float pose[6] = {rx, ry, rz, 0.f, 0.f, 0.f};
  float T[12];
  pose_to_T(pose, T);
  R[0] = T[0]; R[1] = T[1]; R[2] = T[2]; R[3] = T[4]; R[4] = T[5]; R[5] = T[6]; R[6] = T[8]; R[7] = T[9]; R[8] = T[10];
}

// Eigen's 3x3 inverse (cofactors / determinant), what Affine3f::inverse() does to the linear part
__device__ __forceinline__ void inv3(const float m[9], float r[9]) {
#define M_(i, j) m[(i) * 3 + (j)]
#define COF_(i, j) (M_(((i) + 1) % 3, ((j) + 1) % 3) * M_(((i) + 2) % 3, ((j) + 2) % 3) - M_(((i) + 1) % 3, ((j) + 2) % 3) * M_(((i) + 2) % 3, ((j) + 1) % 3))
  const float c00 = COF_(0, 0), c10 = COF_(1, 0), c20 = COF_(2, 0);
  const float det = (c00 * M_(0, 0) + c10 * M_(1, 0)) + c20 * M_(2, 0);
  const float invdet = 1.f / det;
  r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
  r[3] = COF_(0, 1) * invdet; r[4] = COF_(1, 1) * invdet; r[5] = COF_(2, 1) * invdet;
  r[6] = COF_(0, 2) * invdet; r[7] = COF_(1, 2) * invdet; r[8] = COF_(2, 2) * invdet;
#undef COF_
#undef M_
}

__global__ void __launch_bounds__(256)
dsk_apply_kernel(const unsigned char* __restrict__ raw, int n, int stride, const DeskewConst k,
                 const uint32_t* __restrict__ slot, const uint32_t* __restrict__ flags_total,
                 const int* __restrict__ first_idx, const double* __restrict__ imu_tab, float4* __restrict__ out) {
  __shared__ float sRsInv[9];
  const bool deskew = k.enabled && k.n_imu > 1;
  if (threadIdx.x == 0 && deskew) {
    // the first SURVIVING point fixes transStartInverse (:558-562)
    const int f = *first_idx;
    if (f < n) {
      const RawPoint p0 = load_raw(raw, f, stride);
      float rx, ry, rz, R[9], Ri[9];
      find_rotation(k.t_scan + (double)p0.time, imu_tab, k.n_imu, rx, ry, rz);
      rot_from_rpy(rx, ry, rz, R);
      inv3(R, Ri);
      for (int q = 0; q < 9; ++q) sRsInv[q] = Ri[q];
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slot[i];
  const uint32_t s_next = (i + 1 < n) ? slot[i + 1] : *flags_total;
  if (s_next == s) return;  // not a survivor
  const RawPoint p = load_raw(raw, i, stride);
  float4 o = make_float4(p.x, p.y, p.z, p.intensity);
  if (deskew) {
    float rx, ry, rz, R[9], B[9];
    find_rotation(k.t_scan + (double)p.time, imu_tab, k.n_imu, rx, ry, rz);
    rot_from_rpy(rx, ry, rz, R);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        B[a * 3 + b] = (sRsInv[a * 3 + 0] * R[0 * 3 + b] + sRsInv[a * 3 + 1] * R[1 * 3 + b]) + sRsInv[a * 3 + 2] * R[2 * 3 + b];
    o.x = B[0] * p.x + B[1] * p.y + B[2] * p.z + 0.f;
    o.y = B[3] * p.x + B[4] * p.y + B[5] * p.z + 0.f;
    o.z = B[6] * p.x + B[7] * p.y + B[8] * p.z + 0.f;
  }
  out[s] = o;
}

__global__ void dsk_init_kernel(int* first_idx, int n) { *first_idx = n; }

// imu4_host: n_imu x 4 doubles laid out [time | rotX | rotY | rotZ] (pinned staging by the caller)
int deskew_dev(Ctx* c, const void* d_raw, int n, int stride, double t_scan, const double* imu4_host, int n_imu,
               int enabled, DevBuf& out, int* n_out) {
  *n_out = 0;
  if (n <= 0) return LIOGPU_OK;
  LIOGPU_CUDA_OK(c, c->dsk_flags.reserve(((size_t)n + 4) * sizeof(uint32_t)));
  LIOGPU_CUDA_OK(c, c->misc.reserve(256));
  LIOGPU_CUDA_OK(c, out.reserve((size_t)n * sizeof(float4)));
  LIOGPU_CUDA_OK(c, c->imu_tab.reserve((size_t)(n_imu > 0 ? n_imu : 1) * 4 * sizeof(double)));
  if (n_imu > 0)
    LIOGPU_CUDA_OK(c, cudaMemcpyAsync(c->imu_tab.p, imu4_host, (size_t)n_imu * 4 * sizeof(double),
                                      cudaMemcpyHostToDevice, c->stream));
  DeskewConst k;
  k.n_scan = c->prm.n_scan; k.downsample_rate = c->prm.downsample_rate; k.point_filter_num = c->prm.point_filter_num;
  k.min_front = c->prm.lidar_min_front; k.min_back = c->prm.lidar_min_back;
  k.min_left = c->prm.lidar_min_left; k.min_right = c->prm.lidar_min_right;
  k.max_range = c->prm.lidar_max_range; k.max_intensity = c->prm.lidar_max_intensity;
  k.t_scan = t_scan; k.n_imu = n_imu; k.enabled = enabled;
  uint32_t* flags = c->dsk_flags.as<uint32_t>();
  int* d_first = c->misc.as<int>() + 8;
  uint32_t* d_total = c->misc.as<uint32_t>() + 9;
  dsk_init_kernel<<<1, 1, 0, c->stream>>>(d_first, n);
  dsk_flag_kernel<<<div_up(n, 256), 256, 0, c->stream>>>((const unsigned char*)d_raw, n, stride, k, flags, d_first);
  c->launches += 2;
  LIOGPU_CUDA_OK(c, exclusive_scan_u32(c, flags, flags, n, d_total));
  dsk_apply_kernel<<<div_up(n, 256), 256, 0, c->stream>>>((const unsigned char*)d_raw, n, stride, k, flags, d_total,
                                                          d_first, c->imu_tab.as<double>(), out.as<float4>());
  c->launches++;
  LIOGPU_CUDA_OK(c, cudaGetLastError());
  uint32_t* h_total = reinterpret_cast<uint32_t*>((char*)c->h_pinned + 1280);
  LIOGPU_CUDA_OK(c, cudaMemcpyAsync(h_total, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  LIOGPU_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  *n_out = (int)*h_total;
  return LIOGPU_OK;
}

}  // namespace liogpu
