// liorf_oracle_core.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A dependency-free C++14 restatement of liorf's scan-to-map registration hot path, used ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
// for the CUDA library.  Nothing under lio_slam_b200/ includes, links or calls this.
//
// PARITY UNPINNED: the reference (/root/reference, JiLiBIT/LIO-SLAM) ships no tests, golden vectors
// or fixtures for this path and cannot be built here (needs ROS, PCL, FLANN, Eigen, OpenCV, GTSAM —
// none installed, none vendored, all version-unpinned in src/liorf/CMakeLists.txt:27-34).  Each
// function below cites the reference lines it follows (MO = src/liorf/src/mapOptmization.cpp,
// IP = src/liorf/src/imageProjection.cpp) and, where the arithmetic lives in a third-party library,
// restates that library's published algorithm (PCL 1.10 VoxelGrid / getTransformation, FLANN 1.9.1
// L2_Simple exact k-NN, Eigen 3.3.7 ColPivHouseholderQR, OpenCV 4.x gemm/QR solve/Jacobi eigen/LU
// inverse).  The OpenCV pieces ARE pinned against the real library: tests/test_oracle_cv2.py checks
// them against Python cv2 4.13 (available in this image) and against fixtures it generated.
//
// Build: g++ -O2 -std=c++14 -ffp-contract=off -fopenmp (x86-64 baseline: no FMA, like a stock
// Ubuntu build of the reference).  Every float expression is written in the operation order of the
// cited source so that "bit-exact" is well defined.
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace liorf_oracle {

struct P4 {  // packed point: x, y, z, intensity  (pcl::PointXYZI without its padding, UT:65)
  float x, y, z, i;
};

// ---------------------------------------------------------------------------------------------
// a6: pcl::getTransformation(x,y,z,roll,pitch,yaw) as used by trans2Affine3f (MO:887-890) and
// transformPointCloud (MO:856).  Rz*Ry*Rx, f32 products.  Canonical trig: f64 sin/cos rounded to
// f32 (glibc sinf/cosf are <1 ulp but not reproducible on a GPU; SURVEY §7 "transform drift").
// T is row-major 3x4.
inline void pose_to_T(const float pose[6], float T[12]) {
  const float roll = pose[0], pitch = pose[1], yaw = pose[2];
  const float A = (float)std::cos((double)yaw), B = (float)std::sin((double)yaw);
  const float C = (float)std::cos((double)pitch), D = (float)std::sin((double)pitch);
  const float E = (float)std::cos((double)roll), F = (float)std::sin((double)roll);
  const float DE = D * E, DF = D * F;
  T[0] = A * C;  T[1] = A * DF - B * E;  T[2] = B * F + A * DE;  T[3] = pose[3];
  T[4] = B * C;  T[5] = A * E + B * DF;  T[6] = B * DE - A * F;  T[7] = pose[4];
  T[8] = -D;     T[9] = C * F;           T[10] = C * E;          T[11] = pose[5];
}

// pointAssociateToMap (MO:841-847) / transformPointCloud body (MO:862-864): r0*x + r1*y + r2*z + t.
inline P4 apply_T(const float T[12], const P4& p) {
  P4 o;
  o.x = T[0] * p.x + T[1] * p.y + T[2] * p.z + T[3];
  o.y = T[4] * p.x + T[5] * p.y + T[6] * p.z + T[7];
  o.z = T[8] * p.x + T[9] * p.y + T[10] * p.z + T[11];
  o.i = p.i;
  return o;
}

// ---------------------------------------------------------------------------------------------
// a3/a4: pcl::VoxelGrid<PointXYZI>::applyFilter (call sites MO:1536,1582,1609); SURVEY A.1.
// Canonical within-voxel order = ascending input index (stable sort).  Returns 1 when the overflow
// guard fired (output = input), else 0.
inline int voxel_grid(const P4* in, int n, float leaf, std::vector<P4>& out) {
  out.clear();
  if (n <= 0) return 0;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; ++i) {
    const P4& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  const float inv = 1.0f / leaf;
  int64_t d[3];
  for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
  if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) {  // overflow guard (q4)
    out.assign(in, in + n);
    return 1;
  }
  int minb[3], maxb[3], div[3];
  for (int a = 0; a < 3; ++a) {
    minb[a] = (int)std::floor(mn[a] * inv);
    maxb[a] = (int)std::floor(mx[a] * inv);
    div[a] = maxb[a] - minb[a] + 1;
  }
  const int mul1 = div[0], mul2 = div[0] * div[1];
  std::vector<std::pair<uint32_t, int>> keys;
  keys.reserve(n);
  for (int i = 0; i < n; ++i) {
    const P4& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const int ix = (int)(std::floor(p.x * inv) - (float)minb[0]);
    const int iy = (int)(std::floor(p.y * inv) - (float)minb[1]);
    const int iz = (int)(std::floor(p.z * inv) - (float)minb[2]);
    keys.emplace_back((uint32_t)(ix + iy * mul1 + iz * mul2), i);
  }
  std::stable_sort(keys.begin(), keys.end(),
                   [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) {
                     return a.first < b.first;
                   });
  size_t s = 0;
  while (s < keys.size()) {
    size_t e = s;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (e < keys.size() && keys[e].first == keys[s].first) {
      const P4& p = in[keys[e].second];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      ++e;
    }
    const float cnt = (float)(e - s);
    out.push_back(P4{sx / cnt, sy / cnt, sz / cnt, si / cnt});
    s = e;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// a7 (k-NN part): pcl::KdTreeFLANN::nearestKSearch(pointSel, 5, ...) (MO:1631); SURVEY A.2.
// Distance = FLANN L2_Simple<float>: r = 0; r += d*d for x,y,z.  Canonical order: ascending
// (d2, map index).  top[5] is the 6th-best, kept only to log equidistant ties.
struct Knn6 {
  float d2[6];
  int idx[6];
  inline void init() {
    for (int k = 0; k < 6; ++k) { d2[k] = FLT_MAX; idx[k] = INT_MAX; }
  }
  inline bool better_than_worst(float d, int id) const {
    return d < d2[5] || (d == d2[5] && id < idx[5]);
  }
  inline void insert(float d, int id) {
    if (!better_than_worst(d, id)) return;
    int k = 5;
    while (k > 0 && (d < d2[k - 1] || (d == d2[k - 1] && id < idx[k - 1]))) {
      d2[k] = d2[k - 1]; idx[k] = idx[k - 1];
      --k;
    }
    d2[k] = d; idx[k] = id;
  }
  inline bool tie() const {
    return d2[0] == d2[1] || d2[1] == d2[2] || d2[2] == d2[3] || d2[3] == d2[4] || d2[4] == d2[5];
  }
  inline float worst() const { return d2[5]; }
};

// The k smallest L2_Simple distances only (ascending) — what pcl::StatisticalOutlierRemoval reads from
// nearestKSearch(index, meanK + 1, ...) (f2, MO:2513-2514).  Equidistant neighbours carry equal values, so no
// tie rule is involved.
struct KnnDist {
  int k = 0;        // capacity, <= 32
  int found = 0;    // min(k, points inserted)
  float d2[32];
  inline void init() {
    found = 0;
    for (int j = 0; j < 32; ++j) d2[j] = FLT_MAX;
  }
  inline float worst() const { return d2[k - 1]; }
  inline void insert(float d, int /*id*/) {
    if (found < k) ++found;
    if (!(d < d2[k - 1])) return;
    int j = k - 1;
    while (j > 0 && d < d2[j - 1]) { d2[j] = d2[j - 1]; --j; }
    d2[j] = d;
  }
};

inline float l2_simple(const P4& q, const P4& p) {
  float r = 0.f;
  float d = q.x - p.x; r += d * d;
  d = q.y - p.y;       r += d * d;
  d = q.z - p.z;       r += d * d;
  return r;
}

inline void knn_brute(const P4* map, int nm, const P4& q, Knn6& r) {
  r.init();
  for (int j = 0; j < nm; ++j) r.insert(l2_simple(q, map[j]), j);
}

// An exact KD-tree of the FLANN KDTreeSingleIndex family (leaf_max_size 15, split on the widest
// dimension at the median) so the timed CPU baseline has the reference's algorithmic shape (a5:
// kdtreeSurfFromMap->setInputCloud, MO:1846).  Bounds are evaluated in f64 with a relative slack
// so pruning can never drop a point whose f32 L2_Simple distance ties or beats the current worst.
struct KdTree {
  struct Node {
    int left, right;  // children, or -1
    int lo, hi;       // leaf: range in perm
    int dim;
    float cut_lo, cut_hi;
  };
  const P4* pts = nullptr;
  int n = 0;
  std::vector<int> perm;
  std::vector<P4> reord;  // points in perm order (leaf scans are contiguous, like FLANN's reorder)
  std::vector<Node> nodes;
  float bb_lo[3], bb_hi[3];

  static inline float coord(const P4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

  void build(const P4* p, int count) {
    pts = p; n = count;
    perm.resize(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    for (int d = 0; d < 3; ++d) { bb_lo[d] = FLT_MAX; bb_hi[d] = -FLT_MAX; }
    for (int i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) {
        bb_lo[d] = std::min(bb_lo[d], coord(p[i], d));
        bb_hi[d] = std::max(bb_hi[d], coord(p[i], d));
      }
    if (n > 0) build_rec(0, n);
    reord.resize(n);
    for (int i = 0; i < n; ++i) reord[i] = pts[perm[i]];
  }
  int build_rec(int lo, int hi) {
    const int id = (int)nodes.size();
    nodes.push_back(Node{-1, -1, lo, hi, 0, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; ++i)
      for (int d = 0; d < 3; ++d) {
        const float c = coord(pts[perm[i]], d);
        mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c);
      }
    int dim = 0;
    if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
    if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
    const int mid = (lo + hi) / 2;
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return coord(pts[a], dim) < coord(pts[b], dim); });
    float lmax = -FLT_MAX, rmin = FLT_MAX;
    for (int i = lo; i < mid; ++i) lmax = std::max(lmax, coord(pts[perm[i]], dim));
    for (int i = mid; i < hi; ++i) rmin = std::min(rmin, coord(pts[perm[i]], dim));
    const int l = build_rec(lo, mid);
    const int r = build_rec(mid, hi);
    nodes[id].left = l; nodes[id].right = r; nodes[id].dim = dim;
    nodes[id].cut_lo = lmax; nodes[id].cut_hi = rmin;
    return id;
  }
  template <class R>
  void search_rec(int id, const P4& q, R& r, double mind2, double off[3]) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) r.insert(l2_simple(q, reord[i]), perm[i]);
      return;
    }
    const int d = nd.dim;
    const double v = coord(q, d);
    const double d1 = v - nd.cut_lo, d2 = v - nd.cut_hi;
    int best, other;
    double cut;
    if (d1 + d2 < 0) { best = nd.left; other = nd.right; cut = d2 * d2; }
    else { best = nd.right; other = nd.left; cut = d1 * d1; }
    search_rec(best, q, r, mind2, off);
    const double save = off[d];
    const double m2 = mind2 + cut - save;
    if (m2 * 0.999999 <= (double)r.worst()) {
      off[d] = cut;
      search_rec(other, q, r, m2, off);
      off[d] = save;
    }
  }
  template <class R>
  void knn(const P4& q, R& r) const {
    r.init();
    if (n == 0) return;
    double off[3], mind2 = 0;
    for (int d = 0; d < 3; ++d) {
      const double v = coord(q, d);
      double o = 0;
      if (v < bb_lo[d]) o = (v - bb_lo[d]) * (v - bb_lo[d]);
      if (v > bb_hi[d]) o = (v - bb_hi[d]) * (v - bb_hi[d]);
      off[d] = o; mind2 += o;
    }
    search_rec(0, q, r, mind2, off);
  }
};

// ---------------------------------------------------------------------------------------------
// a7 (plane fit): Eigen::Matrix<float,5,3>::colPivHouseholderQr().solve(b) with b = -1 (MO:1633-1648).
// Unblocked column-pivoted Householder QR as in Eigen 3.3.7 ColPivHouseholderQR::computeInPlace +
// _solve_impl; reductions summed left to right; SURVEY A.3.  A is row-major 5x3 and is destroyed.
inline void qr53_solve(float A[5][3], const float b[5], float x[3]) {
  const int rows = 5, cols = 3;
  float hc[3], normU[3], normD[3];
  int perm[3] = {0, 1, 2};
  for (int j = 0; j < cols; ++j) {
    float s = 0.f;
    for (int i = 0; i < rows; ++i) s += A[i][j] * A[i][j];
    normD[j] = normU[j] = std::sqrt(s);
  }
  float maxn = normU[0];
  if (normU[1] > maxn) maxn = normU[1];
  if (normU[2] > maxn) maxn = normU[2];
  const float th = (maxn * FLT_EPSILON) * (maxn * FLT_EPSILON) / (float)rows;
  const float downdate_th = std::sqrt(FLT_EPSILON);
  int nonzero = cols;
  for (int k = 0; k < cols; ++k) {
    int big = k;
    for (int j = k + 1; j < cols; ++j)
      if (normU[j] > normU[big]) big = j;
    const float bigsq = normU[big] * normU[big];
    if (nonzero == cols && bigsq < th * (float)(rows - k)) nonzero = k;
    if (big != k) {
      for (int i = 0; i < rows; ++i) std::swap(A[i][k], A[i][big]);
      std::swap(normU[k], normU[big]);
      std::swap(normD[k], normD[big]);
      std::swap(perm[k], perm[big]);
    }
    // makeHouseholderInPlace on A[k..4][k]
    float tail = 0.f;
    for (int i = k + 1; i < rows; ++i) tail += A[i][k] * A[i][k];
    const float c0 = A[k][k];
    float beta, tau;
    if (tail <= FLT_MIN) {
      tau = 0.f; beta = c0;
      for (int i = k + 1; i < rows; ++i) A[i][k] = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0.f) beta = -beta;
      const float den = c0 - beta;
      for (int i = k + 1; i < rows; ++i) A[i][k] = A[i][k] / den;
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    hc[k] = tau;
    // applyHouseholderOnTheLeft to the trailing columns
    if (tau != 0.f) {
      for (int j = k + 1; j < cols; ++j) {
        float tmp = 0.f;
        for (int i = k + 1; i < rows; ++i) tmp += A[i][k] * A[i][j];
        tmp += A[k][j];
        A[k][j] -= tau * tmp;
        for (int i = k + 1; i < rows; ++i) A[i][j] -= (tau * A[i][k]) * tmp;
      }
    }
    // LAPACK-style norm down-dating
    for (int j = k + 1; j < cols; ++j) {
      if (normU[j] != 0.f) {
        float temp = std::fabs(A[k][j]) / normU[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        const float ratio = normU[j] / normD[j];
        const float temp2 = temp * (ratio * ratio);
        if (temp2 <= downdate_th) {
          float s = 0.f;
          for (int i = k + 1; i < rows; ++i) s += A[i][j] * A[i][j];
          normD[j] = std::sqrt(s);
          normU[j] = normD[j];
        } else {
          normU[j] *= std::sqrt(temp);
        }
      }
    }
  }
  x[0] = x[1] = x[2] = 0.f;
  if (nonzero == 0) return;
  float c[5];
  for (int i = 0; i < rows; ++i) c[i] = b[i];
  for (int k = 0; k < nonzero; ++k) {  // c = H_k ... H_0 applied in order (Q^T b)
    if (hc[k] == 0.f) continue;
    float tmp = 0.f;
    for (int i = k + 1; i < rows; ++i) tmp += A[i][k] * c[i];
    tmp += c[k];
    c[k] -= hc[k] * tmp;
    for (int i = k + 1; i < rows; ++i) c[i] -= (hc[k] * A[i][k]) * tmp;
  }
  for (int i = nonzero - 1; i >= 0; --i) {  // column-oriented back substitution (col-major Eigen)
    c[i] = c[i] / A[i][i];
    for (int r = 0; r < i; ++r) c[r] -= c[i] * A[r][i];
  }
  for (int i = 0; i < nonzero; ++i) x[perm[i]] = c[i];
}

// a7 (per point): body of the OpenMP loop in surfOptimization (MO:1623-1686) after the k-NN.
// nb = the 5 neighbours' coordinates in ascending (d2, idx) order, d2_4 = pointSearchSqDis[4].
// Returns the flag; coeff = (s*pa, s*pb, s*pc, s*pd2).
inline bool plane_residual(const P4& ori, const P4& sel, const P4 nb[5], float d2_4, P4& coeff) {
  coeff = P4{0.f, 0.f, 0.f, 0.f};
  if (!(d2_4 < 1.0)) return false;  // MO:1641
  float A[5][3], b[5], x[3];
  for (int j = 0; j < 5; ++j) {
    A[j][0] = nb[j].x; A[j][1] = nb[j].y; A[j][2] = nb[j].z;
    b[j] = -1.f;
  }
  qr53_solve(A, b, x);
  float pa = x[0], pb = x[1], pc = x[2], pd = 1.f;
  const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);  // MO:1655
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  for (int j = 0; j < 5; ++j) {  // MO:1658-1666 (fabs(float) > 0.2 compares in double)
    const float r = pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd;
    if ((double)std::fabs(r) > 0.2) return false;
  }
  const float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;  // MO:1669
  // MO:1671-1672: the 0.9 literal promotes the quotient to double; sqrt(sqrt(float)) stays float
  const float r2 = ori.x * ori.x + ori.y * ori.y + ori.z * ori.z;
  const float s = (float)(1.0 - 0.9 * (double)std::fabs(pd2) / (double)std::sqrt(std::sqrt(r2)));
  if (!((double)s > 0.1)) return false;  // MO:1679 (coeffSelSurfVec[i] is only written when accepted)
  coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.i = s * pd2;
  return true;
}

// ---------------------------------------------------------------------------------------------
// OpenCV 4.x pieces of LMOptimization (MO:1781-1814); SURVEY A.4.  All row-major 6x6 f32.

// cv::solve(A, b, x, DECOMP_QR) -> hal::QR32f -> QRImpl<float> (Householder, no pivoting), n=6, k=1.
inline bool cv_solve6_qr(const float Ain[36], const float bin[6], float x[6]) {
  const int m = 6, n = 6;
  const float eps = FLT_EPSILON * 10;
  float A[36], b[6], vl[6], hf[6];
  std::memcpy(A, Ain, sizeof(A));
  std::memcpy(b, bin, sizeof(b));
  for (int l = 0; l < n; ++l) {
    const int vs = m - l;
    float vn = 0.f;
    for (int i = 0; i < vs; ++i) { vl[i] = A[(l + i) * 6 + l]; vn += vl[i] * vl[i]; }
    const float t0 = vl[0];
    vl[0] = vl[0] + (vl[0] >= 0.f ? 1.f : -1.f) * std::sqrt(vn);
    vn = std::sqrt(vn + vl[0] * vl[0] - t0 * t0);
    for (int i = 0; i < vs; ++i) vl[i] /= vn;
    for (int j = l; j < n; ++j) {
      float va = 0.f;
      for (int i = l; i < m; ++i) va += vl[i - l] * A[i * 6 + j];
      for (int i = l; i < m; ++i) A[i * 6 + j] -= 2 * vl[i - l] * va;
    }
    hf[l] = vl[0] * vl[0];
    for (int i = 1; i < vs; ++i) A[(l + i) * 6 + l] = vl[i] / vl[0];
  }
  for (int l = 0; l < n; ++l) {
    vl[0] = 1.f;
    for (int j = 1; j < m - l; ++j) vl[j] = A[(j + l) * 6 + l];
    float vb = 0.f;
    for (int i = l; i < m; ++i) vb += vl[i - l] * b[i];
    for (int i = l; i < m; ++i) b[i] -= 2 * vl[i - l] * vb * hf[l];
  }
  for (int i = n - 1; i >= 0; --i) {
    for (int j = n - 1; j > i; --j) b[i] -= b[j] * A[i * 6 + j];
    if (std::fabs(A[i * 6 + i]) < eps) { std::memset(x, 0, 6 * sizeof(float)); return false; }
    b[i] /= A[i * 6 + i];
  }
  std::memcpy(x, b, sizeof(b));
  return true;
}

inline float cv_hypot(float a, float b) {
  a = std::fabs(a); b = std::fabs(b);
  if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
  return 0.f;
}

// cv::eigen(A, E, V) for a symmetric CV_32F matrix -> JacobiImpl_<float>: largest off-diagonal pivot,
// eigenvalues descending in W, eigenvectors as ROWS of V.
inline void cv_eigen6(const float Ain[36], float W[6], float V[36]) {
  const int n = 6;
  const float eps = FLT_EPSILON;
  float A[36];
  std::memcpy(A, Ain, sizeof(A));
  int indR[6], indC[6];
  int i, j, k, m;
  for (i = 0; i < n; ++i) { for (j = 0; j < n; ++j) V[i * 6 + j] = 0.f; V[i * 6 + i] = 1.f; }
  float mv = 0.f;
  for (k = 0; k < n; ++k) {
    W[k] = A[7 * k];
    if (k < n - 1) {
      for (m = k + 1, mv = std::fabs(A[6 * k + m]), i = k + 2; i < n; ++i) {
        const float val = std::fabs(A[6 * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = std::fabs(A[k]), i = 1; i < k; ++i) {
        const float val = std::fabs(A[6 * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  const int maxIters = n * n * 30;
  for (int it = 0; it < maxIters; ++it) {
    for (k = 0, mv = std::fabs(A[indR[0]]), i = 1; i < n - 1; ++i) {
      const float val = std::fabs(A[6 * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < n; ++i) {
      const float val = std::fabs(A[6 * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    const float p = A[6 * k + l];
    if (std::fabs(p) <= eps) break;
    const float y = (float)((W[l] - W[k]) * 0.5);
    float t = std::fabs(y) + cv_hypot(p, y);
    float s = cv_hypot(p, t);
    const float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
    A[6 * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    float a0, b0;
#define LIORF_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (i = 0; i < k; ++i) LIORF_ROT(A[6 * i + k], A[6 * i + l]);
    for (i = k + 1; i < l; ++i) LIORF_ROT(A[6 * k + i], A[6 * i + l]);
    for (i = l + 1; i < n; ++i) LIORF_ROT(A[6 * k + i], A[6 * l + i]);
    for (i = 0; i < n; ++i) LIORF_ROT(V[6 * k + i], V[6 * l + i]);
#undef LIORF_ROT
    for (j = 0; j < 2; ++j) {
      const int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        for (m = idx + 1, mv = std::fabs(A[6 * idx + m]), i = idx + 2; i < n; ++i) {
          const float val = std::fabs(A[6 * idx + i]);
          if (mv < val) mv = val, m = i;
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = std::fabs(A[idx]), i = 1; i < idx; ++i) {
          const float val = std::fabs(A[6 * i + idx]);
          if (mv < val) mv = val, m = i;
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
      std::swap(W[m], W[k]);
      for (i = 0; i < n; ++i) std::swap(V[6 * m + i], V[6 * k + i]);
    }
  }
}

// Mat::inv() (DECOMP_LU) -> hal::LU32f -> LUImpl<float> on [A | I]: partial pivoting, eps = 10*FLT_EPSILON.
inline bool cv_inv6(const float Ain[36], float inv[36]) {
  const int m = 6;
  const float eps = FLT_EPSILON * 10;
  float A[36];
  std::memcpy(A, Ain, sizeof(A));
  float* b = inv;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) b[i * 6 + j] = i == j ? 1.f : 0.f;
  for (int i = 0; i < m; ++i) {
    int k = i;
    for (int j = i + 1; j < m; ++j)
      if (std::fabs(A[j * 6 + i]) > std::fabs(A[k * 6 + i])) k = j;
    if (std::fabs(A[k * 6 + i]) < eps) { std::memset(inv, 0, 36 * sizeof(float)); return false; }
    if (k != i) {
      for (int j = i; j < m; ++j) std::swap(A[i * 6 + j], A[k * 6 + j]);
      for (int j = 0; j < m; ++j) std::swap(b[i * 6 + j], b[k * 6 + j]);
    }
    const float d = -1 / A[i * 6 + i];
    for (int j = i + 1; j < m; ++j) {
      const float alpha = A[j * 6 + i] * d;
      for (int c = i + 1; c < m; ++c) A[j * 6 + c] += alpha * A[i * 6 + c];
      for (int c = 0; c < m; ++c) b[j * 6 + c] += alpha * b[i * 6 + c];
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

// Mat * Mat for CV_32F (cv::gemm, small-matrix path GEMMSingleMul<float,double>): f64 accumulation
// of exact f32 products, rounded once to f32.
inline void cv_gemm6(const float A[36], const float B[36], float C[36]) {
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int k = 0; k < 6; ++k) s += (double)A[i * 6 + k] * (double)B[k * 6 + j];
      C[i * 6 + j] = (float)s;
    }
}
inline void cv_gemv6(const float A[36], const float x[6], float y[6]) {
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += (double)A[i * 6 + k] * (double)x[k];
    y[i] = (float)s;
  }
}

// ---------------------------------------------------------------------------------------------
// a9: Jacobian row of LMOptimization (MO:1714-1778) for one accepted correspondence.
struct Trig { float srx, crx, sry, cry, srz, crz; };
inline Trig lm_trig(const float pose[6]) {  // MO:1714-1719 (canonical f64 trig rounded to f32)
  Trig t;
  t.srx = (float)std::sin((double)pose[2]); t.crx = (float)std::cos((double)pose[2]);
  t.sry = (float)std::sin((double)pose[1]); t.cry = (float)std::cos((double)pose[1]);
  t.srz = (float)std::sin((double)pose[0]); t.crz = (float)std::cos((double)pose[0]);
  return t;
}
inline void jacobian_row(const Trig& g, const P4& o, const P4& c, float row[6], float& rhs) {
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
  rhs = -c.i;
}

// a9: everything in LMOptimization after AtA/AtB exist (MO:1784-1835).  Returns true if converged.
struct LmState {
  float pose[6];
  float matP[36];
  int degenerate;
};
inline bool lm_solve_update(const double JtJ[36], const double Jtr[6], int iterCount, LmState& st,
                            float& deltaR, float& deltaT) {
  float AtA[36], AtB[6], X[6];
  for (int i = 0; i < 36; ++i) AtA[i] = (float)JtJ[i];
  for (int i = 0; i < 6; ++i) AtB[i] = (float)Jtr[i];
  cv_solve6_qr(AtA, AtB, X);
  if (iterCount == 0) {
    float E[6], V[36], V2[36], Vi[36];
    cv_eigen6(AtA, E, V);
    std::memcpy(V2, V, sizeof(V));
    st.degenerate = 0;
    for (int i = 5; i >= 0; --i) {
      if (E[i] < 100.f) {
        for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0.f;
        st.degenerate = 1;
      } else break;
    }
    cv_inv6(V, Vi);
    cv_gemm6(Vi, V2, st.matP);
  }
  if (st.degenerate) {
    float X2[6];
    std::memcpy(X2, X, sizeof(X));
    cv_gemv6(st.matP, X2, X);
  }
  for (int i = 0; i < 6; ++i) st.pose[i] += X[i];
  // MO:1824-1831: pcl::rad2deg(float) is a float multiply by 180/M_PI evaluated... see note in DESIGN
  const double r2d = 180.0 / M_PI;
  const float rx = (float)(X[0] * (float)r2d), ry = (float)(X[1] * (float)r2d), rz = (float)(X[2] * (float)r2d);
  deltaR = (float)std::sqrt((double)rx * rx + (double)ry * ry + (double)rz * rz);
  const float tx = X[3] * 100, ty = X[4] * 100, tz = X[5] * 100;
  deltaT = (float)std::sqrt((double)tx * tx + (double)ty * ty + (double)tz * tz);
  return (double)deltaR < 0.05 && (double)deltaT < 0.05;
}

// ---------------------------------------------------------------------------------------------
// f2: mapOptimization::publishLocalMap (MO:2442-2541) — runs after every registration (MO:504).
// Third-party arithmetic restated here ([3P], versions unpinned like everything else in this file):
//   * Eigen::AngleAxisf(-yaw, UnitZ) -> toRotationMatrix (Eigen 3.3 AngleAxis.h): diagonal
//     cos1_axis .* axis + c, so zz = (1 - c) * 1 * 1 + c; Affine3f::rotate multiplies the (identity)
//     linear part on the right, which changes no value.
//   * pcl::transformPointCloud(in, out, Affine3f) (PCL 1.10 common/impl/transforms.hpp,
//     detail::Transformer<float>::se3, SSE form): per coordinate x*c0 + (y*c1 + (z*c2 + c3)).
//   * pcl::PassThrough (filters/impl/passthrough.hpp): drops non-finite points, keeps min <= v <= max,
//     input order.
//   * pcl::StatisticalOutlierRemoval::applyFilterIndices (filters/impl/statistical_outlier_removal.hpp).
struct LocalMapParams {   // utility.h:219-229
  float left, right, front, back;
  int use_removing_outliers, mean_k;
  float stddev_threshold;
  int use_down_sampling;
  float leaf;
};
struct LocalMapInfo {
  int n_concat, n_cropped, n_after_sor, n_out, leaf_overflow, pad;
  double sor_mean, sor_stddev, sor_threshold;
};

// transformMatrix of MO:2474-2488 as a row-major 3x4; pose_now = transformTobeMapped (MO:2249-2254).
// `cos(-thisPoseYaw)` resolves to the float overload (utility.h:61 `using namespace std`), so every
// operation is f32; canonical trig as in pose_to_T.
inline void yaw_frame_T(const float pose_now[6], float m[12]) {
  const float yaw = pose_now[2], X = pose_now[3], Y = pose_now[4], Z = pose_now[5];
  const float nyaw = -yaw;
  const float c = (float)std::cos((double)nyaw), s = (float)std::sin((double)nyaw);
  const float transformedX = X * c - Y * s;   // MO:2474
  const float transformedY = Y * c + X * s;   // MO:2475
  const float transformedZ = Z;               // MO:2476
  // AngleAxisf(angle = -yaw, axis = (0,0,1)).toRotationMatrix()
  const float ax = 0.f, ay = 0.f, az = 1.f;
  const float sx = s * ax, sy = s * ay, sz = s * az;
  const float c1x = (1.0f - c) * ax, c1y = (1.0f - c) * ay, c1z = (1.0f - c) * az;
  float R[9];
  float tmp = c1x * ay; R[1] = tmp - sz; R[3] = tmp + sz;
  tmp = c1x * az;       R[2] = tmp + sy; R[6] = tmp - sy;
  tmp = c1y * az;       R[5] = tmp - sx; R[7] = tmp + sx;
  R[0] = c1x * ax + c; R[4] = c1y * ay + c; R[8] = c1z * az + c;
  // Identity.linear() * R, row by row (Eigen coefficient-based 3x3 product, summed left to right)
  const float I[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m[4 * i + j] = I[3 * i] * R[j] + I[3 * i + 1] * R[3 + j] + I[3 * i + 2] * R[6 + j];
  m[3] = -transformedX; m[7] = -transformedY; m[11] = -transformedZ;   // MO:2487
}

inline P4 pcl_transform_se3(const float m[12], const P4& p) {   // MO:2489
  P4 o;
  o.x = p.x * m[0] + (p.y * m[1] + (p.z * m[2] + m[3]));
  o.y = p.x * m[4] + (p.y * m[5] + (p.z * m[6] + m[7]));
  o.z = p.x * m[8] + (p.y * m[9] + (p.z * m[10] + m[11]));
  o.i = p.i;
  return o;
}

// PassThrough on x then on y (MO:296-302, 2502-2506)
inline void pass_through_xy(const std::vector<P4>& in, float xmin, float xmax, float ymin, float ymax, std::vector<P4>& out) {
  std::vector<P4> fx;
  for (const P4& p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    if (p.x < xmin || p.x > xmax) continue;
    fx.push_back(p);
  }
  out.clear();
  for (const P4& p : fx) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    if (p.y < ymin || p.y > ymax) continue;
    out.push_back(p);
  }
}

// Second half of StatisticalOutlierRemoval::applyFilterIndices: statistics of the per-point mean distances
// (sequential f64 sums of f32 values / f32 squares), threshold, and the kept indices in input order.
inline void sor_select(const std::vector<float>& distances, int valid_distances, double std_mul, std::vector<int>& kept,
                       double& mean, double& stddev, double& threshold) {
  double sum = 0, sq_sum = 0;
  for (const float& distance : distances) {
    sum += distance;
    sq_sum += distance * distance;
  }
  mean = sum / static_cast<double>(valid_distances);
  const double variance = (sq_sum - sum * sum / static_cast<double>(valid_distances)) / (static_cast<double>(valid_distances) - 1);
  stddev = std::sqrt(variance);
  threshold = mean + std_mul * stddev;
  kept.clear();
  for (int i = 0; i < (int)distances.size(); ++i) {
    if (distances[i] > threshold) continue;   // negative_ == false
    kept.push_back(i);
  }
}

// First half: distances[i] = (float)( sum_{k=1..meanK} sqrt(nn_dists[k]) / meanK ), the sum in f64 over the f32
// square roots of the ascending squared distances; nn_dists[0] is the query itself.  `knn(i, KnnDist&)` is the
// exact (meanK+1)-NN provider (brute force or a KD-tree).  A search that returns fewer than meanK+1 points
// (cloud too small) yields distance 0 and is not counted as valid.
template <class KnnFn>
inline int sor_mean_distances(int n, int mean_k, KnnFn knn, std::vector<float>& distances, int threads) {
  distances.assign(n, 0.f);
  int valid = 0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : valid)
  for (int i = 0; i < n; ++i) {
    KnnDist r;
    r.k = mean_k + 1;
    r.init();
    knn(i, r);
    if (r.found != mean_k + 1) { distances[i] = 0.f; continue; }
    double dist_sum = 0.0;
    for (int k = 1; k < mean_k + 1; ++k) dist_sum += std::sqrt(r.d2[k]);   // float sqrt
    distances[i] = static_cast<float>(dist_sum / mean_k);
    ++valid;
  }
  return valid;
}

// ---------------------------------------------------------------------------------------------
// f3: the loop-closure registration pcl::IterativeClosestPoint<PointXYZI,PointXYZI>::align as configured at
// MO:1111-1121 (max correspondence distance 2*historyKeyframeSearchRadius, 100 iterations, transformation epsilon
// 1e-6, Euclidean fitness epsilon 1e-6, no RANSAC) and icp.getFitnessScore() (MO:1123).  Restated [3P] from PCL
// 1.10: registration/impl/icp.hpp (computeTransformation), correspondence_estimation.hpp
// (determineCorrespondences: nearest neighbour, kept when dist^2 <= max_dist^2), transformation_estimation_svd.hpp
// (Eigen::umeyama without scaling), default_convergence_criteria.hpp, registration.hpp (getFitnessScore).
// Canonical arithmetic where Eigen's is not reproducible (vectorised f32 reductions, f32 JacobiSVD): sums of
// coordinates and of products in f64 (products of two floats are exact), means and the 3x3 covariance rounded
// to f32, rotation from a one-sided Jacobi SVD in f64 of that f32 matrix rounded to f32; everything else in the
// operation order of the cited source.  Nearest-neighbour ties: lower target index.
struct IcpParams {
  double max_correspondence_distance;
  int max_iterations;
  double transformation_epsilon;
  double euclidean_fitness_epsilon;
};
struct IcpResult {
  float final_transformation[16];   // row-major 4x4
  int iterations, converged, state, n_correspondences;
  double fitness_score, last_mse;
};
enum { ICP_NOT_CONVERGED = 0, ICP_ITERATIONS = 1, ICP_TRANSFORM = 2, ICP_ABS_MSE = 3, ICP_REL_MSE = 4,
       ICP_NO_CORRESPONDENCES = 5 };

// R = U diag(1, 1, d) V^T of the f32 3x3 `sigma` (row-major), d = -1 on the smallest singular value when
// det(U) det(V) < 0 (Eigen::umeyama Eq. 39-40).  One-sided Jacobi (Hestenes) in f64.
inline void umeyama_rotation(const float sigma[9], float R[9]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = (double)sigma[3 * i + j];
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; ++i) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (int i = 0; i < 3; ++i) {
          const double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - sn * aq; A[i][q] = sn * ap + c * aq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double sv[3], U[3][3];
  for (int j = 0; j < 3; ++j) sv[j] = std::sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  int lo = 0;
  if (sv[1] < sv[lo]) lo = 1;
  if (sv[2] < sv[lo]) lo = 2;
  const int a = (lo + 1) % 3, b = (lo + 2) % 3;
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) U[i][j] = sv[j] > 0.0 ? A[i][j] / sv[j] : 0.0;
  if (!(sv[lo] > 1e-12 * (sv[a] > sv[b] ? sv[a] : sv[b]))) {  // rank-deficient: complete U right-handed
    U[0][lo] = U[1][a] * U[2][b] - U[2][a] * U[1][b];
    U[1][lo] = U[2][a] * U[0][b] - U[0][a] * U[2][b];
    U[2][lo] = U[0][a] * U[1][b] - U[1][a] * U[0][b];
  }
  auto det3 = [](const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  };
  const double d = det3(U) * det3(V) < 0.0 ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double r = U[i][a] * V[j][a] + U[i][b] * V[j][b] + d * (U[i][lo] * V[j][lo]);
      R[3 * i + j] = (float)r;
    }
}

// transformation_ (row-major 4x4 f32) from the f64 sums over the correspondences (Eigen::umeyama, no scaling)
inline void umeyama_from_sums(const double S[16], int n, float T[16]) {
  // S: sum src (3), sum tgt (3), sum tgt_i * src_j (9, row-major [i][j]), [15] unused
  const double inv_n = 1.0 / (double)n;
  float src_mean[3], dst_mean[3], sigma[9];
  for (int a = 0; a < 3; ++a) { src_mean[a] = (float)(S[a] * inv_n); dst_mean[a] = (float)(S[3 + a] * inv_n); }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)   // one_over_n * dst_demean * src_demean^T
      sigma[3 * i + j] = (float)((S[6 + 3 * i + j] - S[3 + i] * S[j] * inv_n) * inv_n);
  float R[9];
  umeyama_rotation(sigma, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
    // Rt.col(m).head(m) = dst_mean; -= R * src_mean  (f32, products summed left to right)
    T[4 * i + 3] = dst_mean[i] - (R[3 * i] * src_mean[0] + R[3 * i + 1] * src_mean[1] + R[3 * i + 2] * src_mean[2]);
  }
  T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

inline void mat4_mul(const float A[16], const float B[16], float C[16]) {  // Eigen Matrix4f product, f32
  float t[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      t[4 * i + j] = A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j] + A[4 * i + 2] * B[8 + j] + A[4 * i + 3] * B[12 + j];
  std::memcpy(C, t, sizeof(t));
}

// DefaultConvergenceCriteria::hasConverged; returns the convergence state
struct IcpCriteria {
  int max_iterations;
  double rotation_threshold, translation_threshold, mse_threshold_relative, mse_threshold_absolute = 1e-12;
  double prev_mse = DBL_MAX;
  int check(int iterations, const float T[16], double cur_mse) {
    if (iterations >= max_iterations) return ICP_ITERATIONS;
    const float tr = T[0] + T[5] + T[10] - 1;
    const double cos_angle = 0.5 * tr;
    const float tsq = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
    const double translation_sqr = tsq;
    if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) return ICP_TRANSFORM;
    if (std::fabs(cur_mse - prev_mse) < mse_threshold_absolute) return ICP_ABS_MSE;
    if (std::fabs(cur_mse - prev_mse) / prev_mse < mse_threshold_relative) return ICP_REL_MSE;
    prev_mse = cur_mse;
    return ICP_NOT_CONVERGED;
  }
};

// nn(point, idx&, d2&): exact nearest neighbour in the target (ties: lower index).
template <class NnFn>
inline void icp_align(const P4* source, int ns, const P4* target, NnFn nn, const IcpParams& prm, IcpResult& res,
                      int threads) {
  std::vector<P4> cur(source, source + ns);
  std::vector<int> idx(ns);
  std::vector<float> d2(ns);
  float finalT[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  IcpCriteria crit;
  crit.max_iterations = prm.max_iterations;
  crit.rotation_threshold = 1.0 - prm.transformation_epsilon;
  crit.translation_threshold = prm.transformation_epsilon;
  crit.mse_threshold_relative = prm.euclidean_fitness_epsilon;
  const double max_d2 = prm.max_correspondence_distance * prm.max_correspondence_distance;
  std::memset(&res, 0, sizeof(res));
  int it = 0, state = ICP_NOT_CONVERGED;
  do {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int i = 0; i < ns; ++i) nn(cur[i], idx[i], d2[i]);
    double S[16] = {0};
    double mse = 0;
    int cnt = 0;
    for (int i = 0; i < ns; ++i) {
      if (idx[i] < 0 || (double)d2[i] > max_d2) continue;
      ++cnt;
      mse += d2[i];
    }
    res.n_correspondences = cnt;
    if (cnt < 3) { state = ICP_NO_CORRESPONDENCES; break; }   // min_number_correspondences_
    mse /= (double)cnt;
    res.last_mse = mse;
    for (int i = 0; i < ns; ++i) {   // sums for umeyama (f64; every product of two floats is exact)
      if (idx[i] < 0 || (double)d2[i] > max_d2) continue;
      const P4& sp = cur[i];
      const P4& tp = target[idx[i]];
      const double sv[3] = {sp.x, sp.y, sp.z}, tv[3] = {tp.x, tp.y, tp.z};
      for (int a = 0; a < 3; ++a) {
        S[a] += sv[a];
        S[3 + a] += tv[a];
        for (int b = 0; b < 3; ++b) S[6 + 3 * a + b] += tv[a] * sv[b];
      }
    }
    float T[16];
    umeyama_from_sums(S, cnt, T);
    for (int i = 0; i < ns; ++i) {   // pcl::transformPointCloud(Matrix4f): x*c0 + (y*c1 + (z*c2 + c3))
      const P4 p = cur[i];
      cur[i].x = p.x * T[0] + (p.y * T[1] + (p.z * T[2] + T[3]));
      cur[i].y = p.x * T[4] + (p.y * T[5] + (p.z * T[6] + T[7]));
      cur[i].z = p.x * T[8] + (p.y * T[9] + (p.z * T[10] + T[11]));
    }
    mat4_mul(T, finalT, finalT);   // final_transformation_ = transformation_ * final_transformation_
    ++it;
    state = crit.check(it, T, mse);
  } while (state == ICP_NOT_CONVERGED);
  res.iterations = it;
  res.state = state;
  res.converged = (state != ICP_NOT_CONVERGED && state != ICP_NO_CORRESPONDENCES) ? 1 : 0;
  std::memcpy(res.final_transformation, finalT, sizeof(finalT));
  // getFitnessScore(): mean squared nearest-neighbour distance of the source moved by final_transformation_
  double fit = 0;
  int nr = 0;
  std::vector<float> fd(ns);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int i = 0; i < ns; ++i) {
    const P4 p = source[i];
    P4 q;
    q.x = p.x * finalT[0] + (p.y * finalT[1] + (p.z * finalT[2] + finalT[3]));
    q.y = p.x * finalT[4] + (p.y * finalT[5] + (p.z * finalT[6] + finalT[7]));
    q.z = p.x * finalT[8] + (p.y * finalT[9] + (p.z * finalT[10] + finalT[11]));
    q.i = p.i;
    int j;
    nn(q, j, fd[i]);
    if (j < 0) fd[i] = -1.f;
  }
  for (int i = 0; i < ns; ++i)
    if (fd[i] >= 0.f) { fit += fd[i]; ++nr; }
  res.fitness_score = nr > 0 ? fit / nr : DBL_MAX;
}

// ---------------------------------------------------------------------------------------------
// f3 (second half): the Scan Context descriptor of a keyframe cloud, SCManager::makeScancontext
// (include/Scancontext.cpp:151-195) with xy2theta (:23-36), and its ring / sector keys (:198-225), computed at
// every keyframe (MO:2156-2166).  No third-party arithmetic except libm: `sqrt` and `atan` are the C double
// functions here (Scancontext.h has no `using namespace std`), the float arguments are widened, the results
// narrowed where the source assigns to float.  desc is row-major [ring][sector] like the MatrixXd it replaces.
// Non-finite points are skipped (the source would index with an undefined angle).  Key means: Eigen's
// row.mean() = sum / size; the sums are exact in f64 for any realistic heights, so their order is immaterial.
constexpr int SC_NUM_RING = 20, SC_NUM_SECTOR = 60;   // Scancontext.h:82-83

inline float sc_xy2theta(const float& _x, const float& _y) {
  if ((_x >= 0) & (_y >= 0)) return (float)((180 / M_PI) * std::atan((double)(_y / _x)));
  if ((_x < 0) & (_y >= 0)) return (float)(180 - ((180 / M_PI) * std::atan((double)(_y / (-_x)))));
  if ((_x < 0) & (_y < 0)) return (float)(180 + ((180 / M_PI) * std::atan((double)(_y / _x))));
  return (float)(360 - ((180 / M_PI) * std::atan((double)((-_y) / _x))));
}

inline void make_scancontext(const P4* scan, int n, double lidar_height, double max_radius, double* desc,
                             double* ringkey, double* sectorkey) {
  const double NO_POINT = -1000;
  for (int k = 0; k < SC_NUM_RING * SC_NUM_SECTOR; ++k) desc[k] = NO_POINT;
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(scan[i].x) || !std::isfinite(scan[i].y) || !std::isfinite(scan[i].z)) continue;
    const float x = scan[i].x, y = scan[i].y;
    const float z = (float)((double)scan[i].z + lidar_height);   // pt.z = z + LIDAR_HEIGHT (:167)
    const float azim_range = (float)std::sqrt((double)(x * x + y * y));
    const float azim_angle = sc_xy2theta(x, y);
    if ((double)azim_range > max_radius) continue;
    int ring_idx = (int)std::ceil(((double)azim_range / max_radius) * SC_NUM_RING);
    ring_idx = std::max(std::min(SC_NUM_RING, ring_idx), 1);
    const double sc = std::ceil(((double)azim_angle / 360.0) * SC_NUM_SECTOR);
    int sctor_idx = std::isnan(sc) ? 0 : (int)sc;                 // x = y = 0: undefined in the source, bin 1 here
    sctor_idx = std::max(std::min(SC_NUM_SECTOR, sctor_idx), 1);
    double& d = desc[(ring_idx - 1) * SC_NUM_SECTOR + (sctor_idx - 1)];
    if (d < (double)z) d = (double)z;
  }
  for (int k = 0; k < SC_NUM_RING * SC_NUM_SECTOR; ++k)
    if (desc[k] == NO_POINT) desc[k] = 0;
  for (int r = 0; r < SC_NUM_RING; ++r) {
    double sum = 0;
    for (int c = 0; c < SC_NUM_SECTOR; ++c) sum += desc[r * SC_NUM_SECTOR + c];
    ringkey[r] = sum / SC_NUM_SECTOR;
  }
  for (int c = 0; c < SC_NUM_SECTOR; ++c) {
    double sum = 0;
    for (int r = 0; r < SC_NUM_RING; ++r) sum += desc[r * SC_NUM_SECTOR + c];
    sectorkey[c] = sum / SC_NUM_RING;
  }
}

// ---------------------------------------------------------------------------------------------
// f4: which keyframes form the local map — extractNearby (MO:1519-1554) and the selection half of extractCloud
// (MO:1558-1565).  key3d = cloudKeyPoses3D (x, y, z, intensity = keyframe index), key_time = cloudKeyPoses6D[].time.
// Third-party semantics restated [3P]: pcl::KdTreeFLANN::radiusSearch = every point with L2_Simple dist^2 strictly
// below (float)(radius * radius) (FLANN RadiusResultSet), sorted by distance (ties: lower index here);
// nearestKSearch(pt, 1) = the nearest key pose (ties: lower index); VoxelGrid as in voxel_grid().
// ids receives, in the order extractCloud concatenates them, the keyframe index of every entry that survives the
// distance guard of MO:1562 (duplicates are kept, as in the reference).
inline float point_distance(const P4& p1, const P4& p2) {   // lib/common_lib.cpp:34-37
  return (float)std::sqrt((double)((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y) + (p1.z - p2.z) * (p1.z - p2.z)));
}
inline void extract_nearby(const P4* key3d, const double* key_time, int n, double time_laser_info_cur, float search_radius,
                           float density_leaf, std::vector<int>& ids, std::vector<P4>& surrounding_ds) {
  ids.clear();
  surrounding_ds.clear();
  if (n <= 0) return;
  const P4& last = key3d[n - 1];
  const float r2 = (float)((double)search_radius * (double)search_radius);
  std::vector<std::pair<float, int>> hits;
  for (int i = 0; i < n; ++i) {
    const float d2 = l2_simple(last, key3d[i]);
    if (d2 < r2) hits.emplace_back(d2, i);
  }
  std::stable_sort(hits.begin(), hits.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
  std::vector<P4> surrounding(hits.size());
  for (size_t k = 0; k < hits.size(); ++k) surrounding[k] = key3d[hits[k].second];
  voxel_grid(surrounding.data(), (int)surrounding.size(), density_leaf, surrounding_ds);   // MO:1535-1536
  for (P4& pt : surrounding_ds) {                                                         // MO:1537-1541
    int best = 0;
    float bd = FLT_MAX;
    for (int i = 0; i < n; ++i) {
      const float d2 = l2_simple(pt, key3d[i]);
      if (d2 < bd) { bd = d2; best = i; }
    }
    pt.i = key3d[best].i;
  }
  for (int i = n - 1; i >= 0; --i) {                                                      // MO:1544-1551
    if (time_laser_info_cur - key_time[i] < 10.0) surrounding_ds.push_back(key3d[i]);
    else break;
  }
  for (const P4& pt : surrounding_ds) {                                                   // MO:1560-1565
    if (point_distance(pt, last) > search_radius) continue;
    ids.push_back((int)pt.i);
  }
}

}  // namespace liorf_oracle
