// vmp_lio.hpp — host-side mirror of the reference's estimator API (C++), above the C ABI.
//
//   vmp::IESKF        <-> kf::IESKF        (ieskf.h:76-110): x(), P(), predict(); update() runs on the device
//   vmp::LIOBuilder   <-> lio::LIOBuilder  (lio_builder.h:56-83): loadConfig(), process(), status, kf, map
//
// IMU initialisation and IMU propagation stay on the host as in the reference (lio_builder.cpp:28-112); the per-point
// motion compensation (lio_builder.cpp:127-152) and the timed region lio_builder.cpp:224-246 are
// one vmp_scan_raw() call: one upload, one CUDA graph launch.
#pragma once
#include <vector>

#include "../../include/vmp_b200.h"
#include "vmp_state.cuh"

namespace vmp {

struct IMUData { V3 acc, gyro; double timestamp; };          // lio::IMUData   commons.h:12-20
struct CloudPoint { float x, y, z, curvature; };               // fields of pcl::PointXYZINormal the path reads
struct Pose { double offset; V3 acc, gyro, vel, pos; M3 rot; };    // lio::Pose      commons.h:30-43
struct SyncPackage {                                           // lio::SyncPackage commons.h:22-28
    std::vector<IMUData> imus;
    std::vector<CloudPoint> cloud;
    double cloud_start_time = 0.0, cloud_end_time = 0.0;
    // optional non-owning view used instead of `cloud` (the C wrapper processes the caller's buffer in place, no copies)
    CloudPoint* ext_cloud = nullptr;
    size_t ext_size = 0;
    CloudPoint* pts() { return ext_cloud ? ext_cloud : cloud.data(); }
    size_t size() const { return ext_cloud ? ext_size : cloud.size(); }
};
enum LIOStatus { IMU_INIT = 0, MAP_INIT = 1, LIO_MAPPING = 2 };    // lio_builder.h:10-16

class IESKF {
public:
    IESKF();
    St& x() { return x_; }
    double* P() { return P_; }                                 // 23 x 23 row-major
    void predict(const V3& acc, const V3& gyro, double dt, const double* Q /*12x12*/);   // ieskf.cpp:101-123
private:
    St x_;
    double P_[529];
};

class LIOBuilder {
public:
    LIOBuilder() = default;
    ~LIOBuilder();
    int loadConfig(const vmp_config& cfg);                     // lio_builder.cpp:5-26  (creates the device map)
    bool initializeImu(std::vector<IMUData>& imus);            // lio_builder.cpp:28-63
    void undistortCloud(SyncPackage& package, bool compensate = true);   // lio_builder.cpp:65-153 (compensate = false: IMU propagation only)
    bool imu_poses_fit(const SyncPackage& package) const;
    void collectImuSteps(SyncPackage& package, std::vector<vmp_imu_step>& steps);   // the IMU loop of undistortCloud without the propagation itself
    int process(SyncPackage& package, vmp_scan_stats* stats);  // lio_builder.cpp:175-248

    IESKF kf;
    vmp_config config;
    LIOStatus status = IMU_INIT;
    vmp_handle map = nullptr;                                  // std::shared_ptr<VoxelMap> map in the reference
    // LIODataGroup (lio_builder.h:43-54)
    IMUData last_imu;
    std::vector<IMUData> imu_cache;
    std::vector<Pose> imu_poses_cache;
    V3 last_acc = zeros<3, 1>(), last_gyro = zeros<3, 1>();
    double last_cloud_end_time = 0.0;
    double gravity_norm = 0.0;
    double Q[144];
    bool device_undistort = true;                              // motion compensation on the device (vmp_scan_raw) once the map exists
    bool device_predict = false;                               // IMU propagation on the device too (vmp_scan_raw_predict): state / P stay resident there
    bool predict_started = false;                              // the device holds last_acc / last_gyro
    vmp_state prior_x{};                                       // what the last process() handed to the device update
    double prior_P[529] = {};
private:
    std::vector<float> xyz_, ds_;
    std::vector<vmp_pose> poses_;
    std::vector<vmp_imu_step> steps_;
};

}  // namespace vmp
