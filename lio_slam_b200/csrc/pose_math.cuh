// pose_math.cuh — device-side small math of the scan-to-map path.  Compiled with -fmad=false: every
// f32 expression keeps the reference's operation order and rounding (an x86-64 baseline build of the
// reference has no FMA contraction), so results can be compared bit for bit with the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <float.h>

namespace liogpu {

// pcl::getTransformation(x,y,z,roll,pitch,yaw) as used by trans2Affine3f (mapOptmization.cpp:887-890)
// and transformPointCloud (:856): Rz*Ry*Rx with f32 products.  Trig: f64 sin/cos rounded to f32 (the
// canonical choice shared with the oracle; DESIGN.md "transform drift").  T row-major 3x4.
__device__ __forceinline__ void pose_to_T(const float* pose, float* T) {
  const float roll = pose[0], pitch = pose[1], yaw = pose[2];
  const float A = (float)cos((double)yaw), B = (float)sin((double)yaw);
  const float C = (float)cos((double)pitch), D = (float)sin((double)pitch);
  const float E = (float)cos((double)roll), F = (float)sin((double)roll);
  const float DE = D * E, DF = D * F;
  T[0] = A * C;  T[1] = A * DF - B * E;  T[2] = B * F + A * DE;  T[3] = pose[3];
  T[4] = B * C;  T[5] = A * E + B * DF;  T[6] = B * DE - A * F;  T[7] = pose[4];
  T[8] = -D;     T[9] = C * F;           T[10] = C * E;          T[11] = pose[5];
}

// pointAssociateToMap (mapOptmization.cpp:841-847)
__device__ __forceinline__ float4 apply_T(const float* T, const float4 p) {
  float4 o;
  o.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
  o.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
  o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
  o.w = p.w;
  return o;
}

// ---------------------------------------------------------------------------------------------
// Eigen::Matrix<float,5,3>::colPivHouseholderQr().solve(-1) (mapOptmization.cpp:1633-1648), the
// unblocked column-pivoted Householder QR of Eigen 3.3.7.  a[r][c] lives in registers:
// every loop is fully unrolled and the dynamic pivot column is resolved with predicated swaps.
__device__ __forceinline__ void swapf(float& a, float& b) { const float t = a; a = b; b = t; }
__device__ __forceinline__ void swapi(int& a, int& b) { const int t = a; a = b; b = t; }

__device__ __forceinline__ void plane_qr53(float a[5][3], float x[3]) {
  float hc[3], nu[3], nd[3];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) s += a[i][j] * a[i][j];
    nd[j] = nu[j] = sqrtf(s);
  }
  float maxn = nu[0];
  if (nu[1] > maxn) maxn = nu[1];
  if (nu[2] > maxn) maxn = nu[2];
  const float th = (maxn * FLT_EPSILON) * (maxn * FLT_EPSILON) / 5.0f;
  const float downdate_th = sqrtf(FLT_EPSILON);
  int nonzero = 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int big = k;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      // nu[big] with dynamic big: resolve through selects so everything stays in registers
      const float cur = (big == 0) ? nu[0] : ((big == 1) ? nu[1] : nu[2]);
      if (nu[j] > cur) big = j;
    }
    const float nb = (big == 0) ? nu[0] : ((big == 1) ? nu[1] : nu[2]);
    const float bigsq = nb * nb;
    if (nonzero == 3 && bigsq < th * (float)(5 - k)) nonzero = k;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (big == j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) swapf(a[i][k], a[i][j]);
        swapf(nu[k], nu[j]);
        swapf(nd[k], nd[j]);
        swapi(perm[k], perm[j]);
      }
    }
    float tail = 0.f;
#pragma unroll
    for (int i = k + 1; i < 5; ++i) tail += a[i][k] * a[i][k];
    const float c0 = a[k][k];
    float beta, tau;
    if (tail <= FLT_MIN) {
      tau = 0.f; beta = c0;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) a[i][k] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tail);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) a[i][k] = a[i][k] / den;
      tau = (beta - c0) / beta;
    }
    a[k][k] = beta;
    hc[k] = tau;
    if (tau != 0.f) {
#pragma unroll
      for (int j = k + 1; j < 3; ++j) {
        float tmp = 0.f;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) tmp += a[i][k] * a[i][j];
        tmp += a[k][j];
        a[k][j] -= tau * tmp;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) a[i][j] -= (tau * a[i][k]) * tmp;
      }
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (nu[j] != 0.f) {
        float temp = fabsf(a[k][j]) / nu[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = nu[j] / nd[j];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= downdate_th) {
          float s = 0.f;
#pragma unroll
          for (int i = k + 1; i < 5; ++i) s += a[i][j] * a[i][j];
          nd[j] = sqrtf(s);
          nu[j] = nd[j];
        } else {
          nu[j] *= sqrtf(temp);
        }
      }
    }
  }
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero == 0) return;
  float c[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};  // matB0.fill(-1), :1638
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k < nonzero && hc[k] != 0.f) {
      float tmp = 0.f;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) tmp += a[i][k] * c[i];
      tmp += c[k];
      c[k] -= hc[k] * tmp;
#pragma unroll
      for (int i = k + 1; i < 5; ++i) c[i] -= (hc[k] * a[i][k]) * tmp;
    }
  }
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < nonzero) {
      c[i] = c[i] / a[i][i];
#pragma unroll
      for (int r = 0; r < i; ++r) c[r] -= c[i] * a[r][i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < nonzero) {
      if (perm[i] == 0) x[0] = c[i];
      else if (perm[i] == 1) x[1] = c[i];
      else x[2] = c[i];
    }
  }
}

// Body of the OpenMP loop of surfOptimization after the k-NN (mapOptmization.cpp:1641-1684).
// nbr = the five neighbours (ascending distance); returns the accept flag, coeff = (s pa, s pb, s pc, s pd2).
__device__ __forceinline__ bool plane_residual(const float4 ori, const float4 sel, const float4 nbr[5], float4& coeff) {
  float a[5][3], x[3];
#pragma unroll
  for (int j = 0; j < 5; ++j) { a[j][0] = nbr[j].x; a[j][1] = nbr[j].y; a[j][2] = nbr[j].z; }
  plane_qr53(a, x);
  float pa = x[0], pb = x[1], pc = x[2], pd = 1.f;
  const float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool valid = true;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float r = pa * nbr[j].x + pb * nbr[j].y + pc * nbr[j].z + pd;
    if ((double)fabsf(r) > 0.2) valid = false;  // :1662 compares against a double literal
  }
  const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
  const float r2 = ori.x * ori.x + ori.y * ori.y + ori.z * ori.z;
  // :1671 — the 0.9 literal makes the quotient f64; sqrt(sqrt(float)) stays f32 (std:: overloads)
  const float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(r2)));
  coeff = make_float4(s * pa, s * pb, s * pc, s * pd2);
  return valid && ((double)s > 0.1);
}

// ---------------------------------------------------------------------------------------------
// LMOptimization (mapOptmization.cpp:1702-1837)
struct LmTrig { float srx, crx, sry, cry, srz, crz; };
__device__ __forceinline__ LmTrig lm_trig(const float* pose) {  // :1714-1719
  LmTrig t;
  t.srx = (float)sin((double)pose[2]); t.crx = (float)cos((double)pose[2]);
  t.sry = (float)sin((double)pose[1]); t.cry = (float)cos((double)pose[1]);
  t.srz = (float)sin((double)pose[0]); t.crz = (float)cos((double)pose[0]);
  return t;
}
// One row of matA and matB (:1760-1778): row = (arz, ary, arx, cx, cy, cz), rhs = -coeff.intensity.
__device__ __forceinline__ void jacobian_row(const LmTrig g, const float4 o, const float4 c, float row[6], float& rhs) {
  const float srx = g.srx, crx = g.crx, sry = g.sry, cry = g.cry, srz = g.srz, crz = g.crz;
  const float arx = (-srx * cry * o.x - (srx * sry * srz + crx * crz) * o.y + (crx * srz - srx * sry * crz) * o.z) * c.x
                  + (crx * cry * o.x - (srx * crz - crx * sry * srz) * o.y + (crx * sry * crz + srx * srz) * o.z) * c.y;
  const float ary = (-crx * sry * o.x + crx * cry * srz * o.y + crx * cry * crz * o.z) * c.x
                  + (-srx * sry * o.x + srx * sry * srz * o.y + srx * cry * crz * o.z) * c.y
                  + (-cry * o.x - sry * srz * o.y - sry * crz * o.z) * c.z;
  const float arz = ((crx * sry * crz + srx * srz) * o.y + (srx * crz - crx * sry * srz) * o.z) * c.x
                  + ((-crx * srz + srx * sry * crz) * o.y + (-srx * sry * srz - crx * crz) * o.z) * c.y
                  + (cry * crz * o.y - cry * srz * o.z) * c.z;
  row[0] = arz; row[1] = ary; row[2] = arx; row[3] = c.x; row[4] = c.y; row[5] = c.z;
  rhs = -c.w;
}

// ---- OpenCV 4.x 6x6 pieces, single thread, row-major ----
// cv::solve(AtA, AtB, X, DECOMP_QR) (:1784): hal::QR32f Householder, eps = 10*FLT_EPSILON.
static __device__ __noinline__ bool solve6_qr(const float* Ain, const float* bin, float* x) {
  const float eps = FLT_EPSILON * 10;
  float A[36], b[6], vl[6], hf[6];
  for (int i = 0; i < 36; ++i) A[i] = Ain[i];
  for (int i = 0; i < 6; ++i) b[i] = bin[i];
  for (int l = 0; l < 6; ++l) {
    const int vs = 6 - l;
    float vn = 0.f;
    for (int i = 0; i < vs; ++i) { vl[i] = A[(l + i) * 6 + l]; vn += vl[i] * vl[i]; }
    const float t0 = vl[0];
    vl[0] = vl[0] + (vl[0] >= 0.f ? 1.f : -1.f) * sqrtf(vn);
    vn = sqrtf(vn + vl[0] * vl[0] - t0 * t0);
    for (int i = 0; i < vs; ++i) vl[i] /= vn;
    for (int j = l; j < 6; ++j) {
      float va = 0.f;
      for (int i = l; i < 6; ++i) va += vl[i - l] * A[i * 6 + j];
      for (int i = l; i < 6; ++i) A[i * 6 + j] -= 2 * vl[i - l] * va;
    }
    hf[l] = vl[0] * vl[0];
    for (int i = 1; i < vs; ++i) A[(l + i) * 6 + l] = vl[i] / vl[0];
  }
  for (int l = 0; l < 6; ++l) {
    vl[0] = 1.f;
    for (int j = 1; j < 6 - l; ++j) vl[j] = A[(j + l) * 6 + l];
    float vb = 0.f;
    for (int i = l; i < 6; ++i) vb += vl[i - l] * b[i];
    for (int i = l; i < 6; ++i) b[i] -= 2 * vl[i - l] * vb * hf[l];
  }
  for (int i = 5; i >= 0; --i) {
    for (int j = 5; j > i; --j) b[i] -= b[j] * A[i * 6 + j];
    if (fabsf(A[i * 6 + i]) < eps) {
      for (int k = 0; k < 6; ++k) x[k] = 0.f;
      return false;
    }
    b[i] /= A[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i) x[i] = b[i];
  return true;
}

__device__ __forceinline__ float cv_hypotf(float a, float b) {
  a = fabsf(a); b = fabsf(b);
  if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
  return 0.f;
}

// cv::eigen(matAtA, matE, matV) (:1792): JacobiImpl_<float>; eigenvalues descending, eigenvectors as rows.
static __device__ __noinline__ void eigen6_jacobi(const float* Ain, float* W, float* V) {
  const int n = 6;
  float A[36];
  int indR[6], indC[6];
  for (int i = 0; i < 36; ++i) A[i] = Ain[i];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) V[i * 6 + j] = 0.f;
    V[i * 6 + i] = 1.f;
  }
  int i, j, k, m;
  float mv = 0.f;
  for (k = 0; k < n; ++k) {
    W[k] = A[7 * k];
    if (k < n - 1) {
      for (m = k + 1, mv = fabsf(A[6 * k + m]), i = k + 2; i < n; ++i) {
        const float val = fabsf(A[6 * k + i]);
        if (mv < val) { mv = val; m = i; }
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = fabsf(A[k]), i = 1; i < k; ++i) {
        const float val = fabsf(A[6 * i + k]);
        if (mv < val) { mv = val; m = i; }
      }
      indC[k] = m;
    }
  }
  for (int it = 0; it < n * n * 30; ++it) {
    for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < n - 1; ++i) {
      const float val = fabsf(A[6 * i + indR[i]]);
      if (mv < val) { mv = val; k = i; }
    }
    int l = indR[k];
    for (i = 1; i < n; ++i) {
      const float val = fabsf(A[6 * indC[i] + i]);
      if (mv < val) { mv = val; k = indC[i]; l = i; }
    }
    const float p = A[6 * k + l];
    if (fabsf(p) <= FLT_EPSILON) break;
    const float y = (float)((W[l] - W[k]) * 0.5);
    float t = fabsf(y) + cv_hypotf(p, y);
    float s = cv_hypotf(p, t);
    const float c = t / s;
    s = p / s;
    t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[6 * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    float a0, b0;
#define LIOGPU_ROT(v0, v1) { a0 = v0; b0 = v1; v0 = a0 * c - b0 * s; v1 = a0 * s + b0 * c; }
    for (i = 0; i < k; ++i) LIOGPU_ROT(A[6 * i + k], A[6 * i + l])
    for (i = k + 1; i < l; ++i) LIOGPU_ROT(A[6 * k + i], A[6 * i + l])
    for (i = l + 1; i < n; ++i) LIOGPU_ROT(A[6 * k + i], A[6 * l + i])
    for (i = 0; i < n; ++i) LIOGPU_ROT(V[6 * k + i], V[6 * l + i])
#undef LIOGPU_ROT
    for (j = 0; j < 2; ++j) {
      const int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        for (m = idx + 1, mv = fabsf(A[6 * idx + m]), i = idx + 2; i < n; ++i) {
          const float val = fabsf(A[6 * idx + i]);
          if (mv < val) { mv = val; m = i; }
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; ++i) {
          const float val = fabsf(A[6 * i + idx]);
          if (mv < val) { mv = val; m = i; }
        }
        indC[idx] = m;
      }
    }
  }
  for (k = 0; k < n - 1; ++k) {
    m = k;
    for (i = k + 1; i < n; ++i)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      const float tw = W[m]; W[m] = W[k]; W[k] = tw;
      for (i = 0; i < n; ++i) { const float tv = V[6 * m + i]; V[6 * m + i] = V[6 * k + i]; V[6 * k + i] = tv; }
    }
  }
}

// matV.inv() (:1807): cv::invert DECOMP_LU -> hal::LU32f on [A | I], partial pivoting.
static __device__ __noinline__ bool inv6_lu(const float* Ain, float* b) {
  const int m = 6;
  const float eps = FLT_EPSILON * 10;
  float A[36];
  for (int i = 0; i < 36; ++i) A[i] = Ain[i];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) b[i * 6 + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j)
      if (fabsf(A[j * 6 + i]) > fabsf(A[k * 6 + i])) k = j;
    if (fabsf(A[k * 6 + i]) < eps) {
      for (int q = 0; q < 36; ++q) b[q] = 0.f;
      return false;
    }
    if (k != i) {
      for (int j = i; j < m; ++j) swapf(A[i * 6 + j], A[k * 6 + j]);
      for (int j = 0; j < m; ++j) swapf(b[i * 6 + j], b[k * 6 + j]);
    }
    const float d = -1 / A[i * 6 + i];
    for (int j = i + 1; j < m; ++j) {
      const float alpha = A[j * 6 + i] * d;
      for (int q = i + 1; q < m; ++q) A[j * 6 + q] += alpha * A[i * 6 + q];
      for (int q = 0; q < m; ++q) b[j * 6 + q] += alpha * b[i * 6 + q];
    }
  }
  for (int i = m - 1; i >= 0; --i)
    for (int j = 0; j < m; ++j) {
      float s = b[i * 6 + j];
      for (int k = i + 1; k < m; ++k) s -= A[i * 6 + k] * b[k * 6 + j];
      b[i * 6 + j] = s / A[i * 6 + i];
    }
  return true;
}

}  // namespace liogpu
