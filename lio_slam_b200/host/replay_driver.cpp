// replay_driver.cpp — a ROS-free stand-in for laserCloudInfoHandler (mapOptmization.cpp:432-506): replays a
// file of sweeps through the mirrored member functions, one sequence per process / GPU (BASELINE configs[4],
// batch offline mapping: independent sequences, no communication).
//
//   replay_driver <sequence.bin> <out_poses.txt> [device] [scan_leaf] [map_leaf]
//
// sequence.bin: int32 n_scans; per scan: float64 time, float32 guess[6] (what updateInitialGuess would
// provide from IMU odometry), int32 n_points, n_points x PointType (32-byte pcl::PointXYZI records).
// The pose graph is not in scope: the optimised pose of a keyframe is the registered pose (odometry only).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "map_optimization_gpu.h"

using namespace liorf_gpu;

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s sequence.bin poses.txt [device] [scan_leaf] [map_leaf] [publish_local_map 0|1] [select_key_poses_on_device 0|1]\n", argv[0]);
    return 2;
  }
  liogpu_params prm;
  liogpu_default_params(&prm);
  prm.device = argc > 3 ? std::atoi(argv[3]) : 0;
  prm.mapping_surf_leaf_size = argc > 4 ? (float)std::atof(argv[4]) : 0.4f;
  prm.surrounding_keyframe_map_leaf_size = argc > 5 ? (float)std::atof(argv[5]) : 0.5f;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 2; }
  FILE* out = std::fopen(argv[2], "w");
  if (!out) { std::perror(argv[2]); return 2; }
  int32_t n_scans = 0;
  if (std::fread(&n_scans, 4, 1, f) != 1) return 2;
  try {
    mapOptimization MO(prm);
    double gpu_ms = 0, wall_ms = 0, lmap_ms = 0;
    int registered = 0, published = 0;
    const bool publish = argc > 6 && std::atoi(argv[6]) != 0;  // also run publishLocalMap after every scan (:504)
    MO.selectKeyPosesOnDevice = argc > 7 && std::atoi(argv[7]) != 0;
    if (publish) {
      MO.localMapKeyFramesNumber = 50;                            // 6t.yaml:17
      MO.localMapParams.local_mapping_surf_leaf_size = 0.2f;      // jeep.yaml:23
    }
    for (int s = 0; s < n_scans; ++s) {
      double t;
      float guess[6];
      int32_t n;
      if (std::fread(&t, 8, 1, f) != 1 || std::fread(guess, 4, 6, f) != 6 || std::fread(&n, 4, 1, f) != 1) return 2;
      MO.laserCloudSurfLast.resize(n);
      if ((int)std::fread(MO.laserCloudSurfLast.data(), sizeof(PointType), n, f) != n) return 2;
      MO.timeLaserInfoCur = t;
      const auto t0 = std::chrono::steady_clock::now();
      std::memcpy(MO.transformTobeMapped, guess, sizeof(guess));  // updateInitialGuess (:1438-1502) stays host
      MO.extractSurroundingKeyFrames();
      MO.downsampleAndScan2Map();   // downsampleCurrentScan + scan2MapOptimization
      if (MO.lastStatus < 0) { std::fprintf(stderr, "scan %d: %s\n", s, MO.lastError()); return 1; }
      if (!MO.cloudKeyPoses3D.empty()) { gpu_ms += MO.lastInfo.gpu_ms; ++registered; }
      if (MO.saveFrame()) MO.saveKeyFrame();
      if (publish) {
        MO.publishLocalMap();
        if (MO.lastStatus < 0) { std::fprintf(stderr, "scan %d publishLocalMap: %s\n", s, MO.lastError()); return 1; }
        lmap_ms += MO.lastLocalMapInfo.gpu_ms;
        ++published;
      }
      wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      std::fprintf(out, "%d %.9g %.9g %.9g %.9g %.9g %.9g %d %d %d %d\n", s, MO.transformTobeMapped[0], MO.transformTobeMapped[1],
                   MO.transformTobeMapped[2], MO.transformTobeMapped[3], MO.transformTobeMapped[4], MO.transformTobeMapped[5],
                   MO.lastInfo.iterations, MO.lastInfo.n_sel, (int)MO.cloudKeyPoses3D.size(), MO.laserCloudSurfFromMapDSNum);
    }
    std::printf("{\"scans\": %d, \"registered\": %d, \"keyframes\": %d, \"wall_ms_per_scan\": %.4f, \"loop_gpu_ms_per_scan\": %.4f, "
                "\"local_map_gpu_ms_per_scan\": %.4f, \"local_map_points\": %d}\n",
                n_scans, registered, (int)MO.cloudKeyPoses3D.size(), wall_ms / n_scans, registered ? gpu_ms / registered : 0.0,
                published ? lmap_ms / published : 0.0, (int)MO.localMapCloud.size());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  std::fclose(f);
  std::fclose(out);
  return 0;
}
