// pcl_pin.cpp — runs the REAL PCL / FLANN / Eigen / OpenCV calls of liorf's scan-to-map path on the committed seeded
// inputs and writes their outputs (LPIN records, see lpin.py) for tests/test_pcl_pin.py to hold the oracle against.
// Build and run: README.md (ROS Noetic: PCL 1.10, Eigen 3.3.7, OpenCV 4.2).  Every block names the reference call site
// it reproduces (src/liorf/src/mapOptmization.cpp unless noted) and uses the same types and setters.
//
//   usage: pcl_pin pin_inputs.lpin out/pcl_pin_outputs.lpin
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/core.hpp>
#include <pcl/common/eigen.h>
#include <pcl/common/transforms.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/statistical_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/icp.h>

typedef pcl::PointXYZI PointType;  // utility.h:65

// ---------------------------------------------------------------- LPIN records
struct Rec {
  int dtype = 0, ndim = 0;  // 0 f32, 1 f64, 2 i32, 3 u8
  int64_t shape[4] = {0, 0, 0, 0};
  std::vector<unsigned char> data;
  size_t count() const { size_t c = 1; for (int i = 0; i < ndim; ++i) c *= (size_t)shape[i]; return c; }
  const float* f32() const { return reinterpret_cast<const float*>(data.data()); }
};
static const int kSize[4] = {4, 8, 4, 1};

static std::map<std::string, Rec> read_lpin(const char* path) {
  std::map<std::string, Rec> out;
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  char magic[8];
  int32_t n = 0;
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "LPIN1\0\0\0", 8) != 0 || std::fread(&n, 4, 1, f) != 1) { std::fprintf(stderr, "bad LPIN file\n"); std::exit(2); }
  for (int r = 0; r < n; ++r) {
    char name[33] = {0};
    Rec rec;
    int32_t hd[2];
    if (std::fread(name, 1, 32, f) != 32 || std::fread(hd, 4, 2, f) != 2 || std::fread(rec.shape, 8, 4, f) != 4) std::exit(2);
    rec.dtype = hd[0]; rec.ndim = hd[1];
    rec.data.resize(rec.count() * (size_t)kSize[rec.dtype]);
    if (std::fread(rec.data.data(), 1, rec.data.size(), f) != rec.data.size()) std::exit(2);
    out[name] = rec;
  }
  std::fclose(f);
  return out;
}

struct Writer {
  std::vector<std::pair<std::string, Rec>> recs;
  template <class T> void add(const std::string& name, int dtype, const std::vector<T>& v, std::vector<int64_t> shape) {
    Rec r;
    r.dtype = dtype; r.ndim = (int)shape.size();
    for (size_t i = 0; i < shape.size(); ++i) r.shape[i] = shape[i];
    r.data.resize(v.size() * sizeof(T));
    std::memcpy(r.data.data(), v.data(), r.data.size());
    recs.emplace_back(name, r);
  }
  void cloud(const std::string& name, const pcl::PointCloud<PointType>& c) {
    std::vector<float> v;
    for (const auto& p : c.points) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); v.push_back(p.intensity); }
    add(name, 0, v, {(int64_t)c.points.size(), 4});
  }
  void save(const char* path) {
    FILE* f = std::fopen(path, "wb");
    if (!f) { std::perror(path); std::exit(2); }
    std::fwrite("LPIN1\0\0\0", 1, 8, f);
    const int32_t n = (int32_t)recs.size();
    std::fwrite(&n, 4, 1, f);
    for (auto& kv : recs) {
      char name[32] = {0};
      std::strncpy(name, kv.first.c_str(), 31);
      const int32_t hd[2] = {kv.second.dtype, kv.second.ndim};
      std::fwrite(name, 1, 32, f); std::fwrite(hd, 4, 2, f); std::fwrite(kv.second.shape, 8, 4, f);
      std::fwrite(kv.second.data.data(), 1, kv.second.data.size(), f);
    }
    std::fclose(f);
  }
};

static pcl::PointCloud<PointType>::Ptr to_cloud(const Rec& r) {
  pcl::PointCloud<PointType>::Ptr c(new pcl::PointCloud<PointType>());
  const float* p = r.f32();
  for (int64_t i = 0; i < r.shape[0]; ++i) {
    PointType q;
    q.x = p[4 * i]; q.y = p[4 * i + 1]; q.z = p[4 * i + 2]; q.intensity = p[4 * i + 3];
    c->push_back(q);
  }
  return c;
}

static pcl::PointCloud<PointType>::Ptr voxel(const pcl::PointCloud<PointType>::Ptr& in, float leaf) {
  pcl::VoxelGrid<PointType> f;  // :156-160
  f.setLeafSize(leaf, leaf, leaf);  // :286-289
  f.setInputCloud(in);
  pcl::PointCloud<PointType>::Ptr out(new pcl::PointCloud<PointType>());
  f.filter(*out);  // :1536 / :1582 / :1609
  return out;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s pin_inputs.lpin outputs.lpin\n", argv[0]); return 2; }
  auto in = read_lpin(argv[1]);
  Writer w;
  auto A = to_cloud(in.at("cloud_a")), B = to_cloud(in.at("cloud_b"));
  const float* g = in.at("pose_guess").f32();  // {roll, pitch, yaw, x, y, z} = transformTobeMapped (:171)

  // ---- VoxelGrid (:1536, :1582, :1609) ----
  auto A04 = voxel(A, 0.4f), B05 = voxel(B, 0.5f), A20 = voxel(A, 2.0f);
  w.cloud("vox_a_04", *A04); w.cloud("vox_b_05", *B05); w.cloud("vox_a_20", *A20);
  {
    auto guard = voxel(A, 0.001f);  // index overflow: PCL warns and returns the input (quirk q4)
    w.add<int32_t>("vox_guard_n", 2, {(int32_t)guard->size(), (int32_t)A->size()}, {2});
  }

  // ---- pcl::getTransformation / trans2Affine3f (:887-890) and pointAssociateToMap (:841-847) ----
  Eigen::Affine3f T = pcl::getTransformation(g[3], g[4], g[5], g[0], g[1], g[2]);
  {
    std::vector<float> m(16);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m[4 * r + c] = T(r, c);
    w.add("T_pose", 0, m, {4, 4});
  }

  // ---- kdtreeSurfFromMap->setInputCloud (:1846) + nearestKSearch (:1631) + plane fit (:1633-1648) ----
  pcl::KdTreeFLANN<PointType> kd;
  kd.setInputCloud(B05);
  std::vector<int32_t> knn_idx;
  std::vector<float> knn_d2, qr_x, sel_xyz;
  for (const auto& pi : A04->points) {
    PointType ps;  // :843-846
    ps.x = T(0, 0) * pi.x + T(0, 1) * pi.y + T(0, 2) * pi.z + T(0, 3);
    ps.y = T(1, 0) * pi.x + T(1, 1) * pi.y + T(1, 2) * pi.z + T(1, 3);
    ps.z = T(2, 0) * pi.x + T(2, 1) * pi.y + T(2, 2) * pi.z + T(2, 3);
    ps.intensity = pi.intensity;
    sel_xyz.push_back(ps.x); sel_xyz.push_back(ps.y); sel_xyz.push_back(ps.z);
    std::vector<int> ind;
    std::vector<float> d2;
    kd.nearestKSearch(ps, 5, ind, d2);
    Eigen::Matrix<float, 5, 3> matA0;
    Eigen::Matrix<float, 5, 1> matB0;
    matA0.setZero();
    matB0.fill(-1);
    for (int j = 0; j < 5; ++j) {
      knn_idx.push_back(j < (int)ind.size() ? ind[j] : -1);
      knn_d2.push_back(j < (int)d2.size() ? d2[j] : -1.f);
      if (j < (int)ind.size()) {
        matA0(j, 0) = B05->points[ind[j]].x; matA0(j, 1) = B05->points[ind[j]].y; matA0(j, 2) = B05->points[ind[j]].z;
      }
    }
    Eigen::Vector3f matX0 = matA0.colPivHouseholderQr().solve(matB0);  // :1648
    qr_x.push_back(matX0(0)); qr_x.push_back(matX0(1)); qr_x.push_back(matX0(2));
  }
  const int64_t nq = (int64_t)A04->size();
  w.add("knn_idx", 2, knn_idx, {nq, 5}); w.add("knn_d2", 0, knn_d2, {nq, 5});
  w.add("qr_x", 0, qr_x, {nq, 3}); w.add("point_sel", 0, sel_xyz, {nq, 3});

  // ---- publishLocalMap (:2474-2516): AngleAxisf, transformPointCloud, PassThrough x / y, StatisticalOutlierRemoval ----
  const float* pn = in.at("pose_now").f32();
  {
    const float yaw = pn[2], X = pn[3], Y = pn[4], Z = pn[5];
    const float tX = X * std::cos(-yaw) - Y * std::sin(-yaw);  // :2474
    const float tY = Y * std::cos(-yaw) + X * std::sin(-yaw);  // :2475
    Eigen::Affine3f t2 = Eigen::Affine3f::Identity();
    t2.translation() << -tX, -tY, -Z;
    t2.rotate(Eigen::AngleAxisf(-yaw, Eigen::Vector3f::UnitZ()));  // :2481-2486
    std::vector<float> R(16);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) R[4 * r + c] = t2(r, c);
    w.add("aa_R", 0, R, {4, 4});
    pcl::PointCloud<PointType>::Ptr moved(new pcl::PointCloud<PointType>());
    pcl::transformPointCloud(*B, *moved, t2);  // :2488
    w.cloud("tpc_out", *moved);
    pcl::PassThrough<PointType> px, py;  // :293-301, utility.h:219-223
    px.setFilterFieldName("x"); px.setFilterLimits(-40.0f, 40.0f);
    py.setFilterFieldName("y"); py.setFilterLimits(-20.0f, 70.0f);
    pcl::PointCloud<PointType>::Ptr c1(new pcl::PointCloud<PointType>()), c2(new pcl::PointCloud<PointType>());
    px.setInputCloud(moved); px.filter(*c1);  // :2502-2504
    py.setInputCloud(c1); py.filter(*c2);     // :2505-2507
    w.cloud("pass_out", *c2);
    pcl::StatisticalOutlierRemoval<PointType> sor;  // :2510-2516
    sor.setInputCloud(c2);
    sor.setMeanK(10);
    sor.setStddevMulThresh(1.0);
    pcl::PointCloud<PointType>::Ptr c3(new pcl::PointCloud<PointType>());
    sor.filter(*c3);
    w.cloud("sor_out", *c3);
  }

  // ---- loop-closure ICP (:1111-1123) ----
  {
    auto src = to_cloud(in.at("icp_source"));
    pcl::IterativeClosestPoint<PointType, PointType> icp;
    icp.setMaxCorrespondenceDistance(10.0 * 2);  // historyKeyframeSearchRadius*2, :1112, utility.h:321
    icp.setMaximumIterations(100);
    icp.setTransformationEpsilon(1e-6);
    icp.setEuclideanFitnessEpsilon(1e-6);
    icp.setRANSACIterations(0);
    icp.setInputSource(src);
    icp.setInputTarget(B05);
    pcl::PointCloud<PointType>::Ptr unused(new pcl::PointCloud<PointType>());
    icp.align(*unused);
    Eigen::Matrix4f Tf = icp.getFinalTransformation();
    std::vector<float> m(16);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m[4 * r + c] = Tf(r, c);
    w.add("icp_T", 0, m, {4, 4});
    w.add<double>("icp_meta", 1, {icp.hasConverged() ? 1.0 : 0.0, icp.getFitnessScore()}, {2});
  }

  // ---- LMOptimization's OpenCV pieces (:1781-1814) on a committed A (n x 6), b (n) ----
  {
    const Rec& rA = in.at("lm_A");
    const Rec& rb = in.at("lm_b");
    const int n = (int)rA.shape[0];
    cv::Mat matA(n, 6, CV_32F), matB(n, 1, CV_32F);
    std::memcpy(matA.data, rA.f32(), (size_t)n * 6 * 4);
    std::memcpy(matB.data, rb.f32(), (size_t)n * 4);
    cv::Mat matAt, matAtA, matAtB, matX(6, 1, CV_32F, cv::Scalar::all(0));
    cv::transpose(matA, matAt);
    matAtA = matAt * matA;  // :1782
    matAtB = matAt * matB;  // :1783
    cv::solve(matAtA, matAtB, matX, cv::DECOMP_QR);  // :1784
    cv::Mat matE(1, 6, CV_32F, cv::Scalar::all(0)), matV(6, 6, CV_32F, cv::Scalar::all(0));
    cv::eigen(matAtA, matE, matV);  // :1792
    cv::Mat Vinv = matV.inv();      // :1807
    auto dump = [&](const char* name, const cv::Mat& m) {
      std::vector<float> v((size_t)m.rows * m.cols);
      for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) v[(size_t)r * m.cols + c] = m.at<float>(r, c);
      w.add(name, 0, v, {m.rows, m.cols});
    };
    dump("cv_AtA", matAtA); dump("cv_AtB", matAtB); dump("cv_x", matX); dump("cv_E", matE); dump("cv_V", matV); dump("cv_Vinv", Vinv);
  }

  // library versions, for the record
  {
    std::vector<int32_t> v = {PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION, EIGEN_WORLD_VERSION, EIGEN_MAJOR_VERSION,
                              EIGEN_MINOR_VERSION, CV_VERSION_MAJOR, CV_VERSION_MINOR, CV_VERSION_REVISION};
    w.add("versions", 2, v, {9});
  }
  w.save(argv[2]);
  std::printf("wrote %s (%zu records)\n", argv[2], w.recs.size());
  return 0;
}
