// vmp_lio.cu — host side of the estimator: IMU init, IMU propagation, undistortion,
// the process() state machine, and the vmp_lio_* C wrappers.  No device code in here; the
// measurement update and the map live behind vmp_scan / vmp_first_scan.
#include "vmp_lio.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vmp {

void set_error(const char* fmt, ...);

namespace {
// non-zero entries of each row of a small dense matrix, ascending column
struct SparseRows { int n[23]; int col[23][12]; double val[23][12]; };
void sparse_rows(const double* M, int rows, int cols, SparseRows& s) {
    for (int i = 0; i < rows; i++) {
        int c = 0;
        for (int j = 0; j < cols; j++) if (M[i * cols + j] != 0.0 && c < 12) { s.col[i][c] = j; s.val[i][c] = M[i * cols + j]; c++; }
        s.n[i] = c;
    }
}
template <int BR, int BC>
void put(double* M, int ld, int r0, int c0, const Mat<BR, BC>& b) { for (int i = 0; i < BR; i++) for (int j = 0; j < BC; j++) M[(r0 + i) * ld + c0 + j] = b(i, j); }
}  // namespace

IESKF::IESKF() {
    x_.pos = zeros<3, 1>(); x_.rot = eye<3>(); x_.rot_ext = eye<3>(); x_.pos_ext = zeros<3, 1>();
    x_.vel = zeros<3, 1>(); x_.bg = zeros<3, 1>(); x_.ba = zeros<3, 1>(); x_.g = v3(0.0, 0.0, -GRAVITY);
    std::memset(P_, 0, sizeof(P_));
}

// ieskf.cpp:101-123
void IESKF::predict(const V3& acc_in, const V3& gyro_in, double dt, const double* Q) {
    const V3 w = sub(gyro_in, x_.bg);
    const V3 a = sub(acc_in, x_.ba);
    const V3 dpos = scale(x_.vel, dt);
    const V3 drot = scale(w, dt);
    const V3 dvel = scale(add(mul(x_.rot, a), x_.g), dt);

    static thread_local double F[529], G[23 * 12], T1[529], T2[529], T3[23 * 12];
    std::memset(F, 0, sizeof(F));
    for (int i = 0; i < 23; i++) F[i * 23 + i] = 1.0;
    put(F, 23, 0, 12, scale(eye<3>(), dt));
    put(F, 23, 3, 3, so3_exp(scale(neg(w), dt)));
    put(F, 23, 3, 15, scale(neg(right_jacobian(scale(w, dt))), dt));
    put(F, 23, 12, 3, scale(mul(neg(x_.rot), hat(a)), dt));
    put(F, 23, 12, 18, scale(neg(x_.rot), dt));
    put(F, 23, 12, 21, scale(st_Mx(x_.g), dt));
    put(F, 23, 21, 21, mul(st_Nx(x_.g), st_Mx(x_.g)));
    std::memset(G, 0, sizeof(G));
    put(G, 12, 3, 0, scale(neg(right_jacobian(scale(w, dt))), dt));
    put(G, 12, 12, 3, scale(neg(x_.rot), dt));
    put(G, 12, 15, 6, scale(eye<3>(), dt));
    put(G, 12, 18, 9, scale(eye<3>(), dt));
    // x_ += delta (Vector24d variant, ieskf.cpp:23-33): only pos, rot, vel are non-zero
    x_.pos = add(x_.pos, dpos);
    x_.rot = mul(x_.rot, so3_exp(drot));
    x_.rot_ext = mul(x_.rot_ext, so3_exp(zeros<3, 1>()));
    x_.pos_ext = add(x_.pos_ext, zeros<3, 1>());
    x_.vel = add(x_.vel, dvel);
    x_.bg = add(x_.bg, zeros<3, 1>());
    x_.ba = add(x_.ba, zeros<3, 1>());
    x_.g = mul(so3_exp(zeros<3, 1>()), x_.g);
    // P = F P F^T + G Q G^T.  F is the identity plus seven small blocks and G has four, so the products skip the exact
    // zeros; the remaining terms are taken in the same ascending-k order as the dense left-to-right product
    // (x + 0*y == x), i.e. the result is the dense one at ~1/8 of the work.
    static thread_local SparseRows Fs, Gs;
    sparse_rows(F, 23, 23, Fs);
    sparse_rows(G, 23, 12, Gs);
    for (int i = 0; i < 23; i++) {                                // T1 = F P, row i as a sum of scaled rows of P (vectorises over j)
        double* t = T1 + i * 23;
        const double v0 = Fs.val[i][0];
        const double* p0 = P_ + Fs.col[i][0] * 23;
        for (int j = 0; j < 23; j++) t[j] = v0 * p0[j];
        for (int e = 1; e < Fs.n[i]; e++) {
            const double v = Fs.val[i][e];
            const double* pe = P_ + Fs.col[i][e] * 23;
            for (int j = 0; j < 23; j++) t[j] += v * pe[j];
        }
    }
    for (int i = 0; i < 23; i++)                                  // T2 = T1 F^T
        for (int j = 0; j < 23; j++) {
            double sacc = T1[i * 23 + Fs.col[j][0]] * Fs.val[j][0];
            for (int e = 1; e < Fs.n[j]; e++) sacc += T1[i * 23 + Fs.col[j][e]] * Fs.val[j][e];
            T2[i * 23 + j] = sacc;
        }
    std::memset(T3, 0, sizeof(T3));
    for (int i = 0; i < 23; i++)                                  // T3 = G Q
        for (int c = 0; c < 12 && Gs.n[i] > 0; c++) {
            double sacc = Gs.val[i][0] * Q[Gs.col[i][0] * 12 + c];
            for (int e = 1; e < Gs.n[i]; e++) sacc += Gs.val[i][e] * Q[Gs.col[i][e] * 12 + c];
            T3[i * 12 + c] = sacc;
        }
    for (int i = 0; i < 23; i++)                                  // P = T2 + T3 G^T
        for (int j = 0; j < 23; j++) {
            double sacc = 0.0;
            if (Gs.n[i] > 0 && Gs.n[j] > 0) {
                sacc = T3[i * 12 + Gs.col[j][0]] * Gs.val[j][0];
                for (int e = 1; e < Gs.n[j]; e++) sacc += T3[i * 12 + Gs.col[j][e]] * Gs.val[j][e];
            }
            P_[i * 23 + j] = T2[i * 23 + j] + sacc;
        }
}

LIOBuilder::~LIOBuilder() { if (map) vmp_destroy(map); }

// lio_builder.cpp:5-26
int LIOBuilder::loadConfig(const vmp_config& cfg) {
    config = cfg;
    status = IMU_INIT;
    std::memset(Q, 0, sizeof(Q));
    for (int i = 0; i < 12; i++) Q[i * 12 + i] = 1.0;
    for (int i = 0; i < 3; i++) {
        Q[i * 12 + i] = cfg.ng; Q[(3 + i) * 12 + 3 + i] = cfg.na;
        Q[(6 + i) * 12 + 6 + i] = cfg.nbg; Q[(9 + i) * 12 + 9 + i] = cfg.nba;
    }
    // scan_filter.setLeafSize (lio_builder.cpp:13-14): the filter runs on the device (vmp_downsample / inside vmp_scan_raw)
    if (map) { vmp_destroy(map); map = nullptr; }
    return vmp_create(&cfg, &map);
}

// lio_builder.cpp:28-63
bool LIOBuilder::initializeImu(std::vector<IMUData>& imus) {
    imu_cache.insert(imu_cache.end(), imus.begin(), imus.end());
    if (imu_cache.size() < (size_t)config.imu_init_num) return false;
    V3 acc_mean = zeros<3, 1>(), gyro_mean = zeros<3, 1>();
    for (const auto& imu : imu_cache) { acc_mean = add(acc_mean, imu.acc); gyro_mean = add(gyro_mean, imu.gyro); }
    acc_mean = divs(acc_mean, static_cast<double>(imu_cache.size()));
    gyro_mean = divs(gyro_mean, static_cast<double>(imu_cache.size()));
    gravity_norm = norm(acc_mean);
    St& x = kf.x();
    std::memcpy(x.rot_ext.a, config.r_il, sizeof(double) * 9);
    std::memcpy(x.pos_ext.a, config.p_il, sizeof(double) * 3);
    x.bg = gyro_mean;
    if (config.gravity_align) {
        x.rot = rot_from_two_vectors(normalized(neg(acc_mean)), v3(0.0, 0.0, -1.0));
        x.g = scale(normalized(v3(0, 0, -1.0)), GRAVITY);
    } else {
        x.g = scale(normalized(neg(acc_mean)), GRAVITY);
    }
    double* P = kf.P();
    std::memset(P, 0, sizeof(double) * 529);
    for (int i = 0; i < 23; i++) P[i * 23 + i] = 1.0;
    for (int i = 0; i < 3; i++) {
        P[(6 + i) * 23 + 6 + i] = 0.00001; P[(9 + i) * 23 + 9 + i] = 0.00001;
        P[(15 + i) * 23 + 15 + i] = 0.0001; P[(18 + i) * 23 + 18 + i] = 0.0001;
    }
    P[21 * 23 + 21] = 0.00001; P[22 * 23 + 22] = 0.00001;
    last_imu = imus.back();
    return true;
}

// lio_builder.cpp:65-153; compensate = false leaves the point loop (:127-152) to the device (vmp_scan_raw)
void LIOBuilder::undistortCloud(SyncPackage& package, bool compensate) {
    imu_cache.clear();
    imu_cache.push_back(last_imu);
    imu_cache.insert(imu_cache.end(), package.imus.begin(), package.imus.end());
    const double imu_time_end = imu_cache.back().timestamp;
    const double cloud_time_begin = package.cloud_start_time, cloud_time_end = package.cloud_end_time;
    if (compensate)
        std::stable_sort(package.pts(), package.pts() + package.size(),
                         [](const CloudPoint& a, const CloudPoint& b) { return a.curvature < b.curvature; });
    imu_poses_cache.clear();
    imu_poses_cache.push_back(Pose{0.0, last_acc, last_gyro, kf.x().vel, kf.x().pos, kf.x().rot});
    V3 acc_val = zeros<3, 1>(), gyro_val = zeros<3, 1>();
    double dt = 0.0;
    for (size_t i = 0; i + 1 < imu_cache.size(); i++) {
        const IMUData& head = imu_cache[i];
        const IMUData& tail = imu_cache[i + 1];
        if (tail.timestamp < last_cloud_end_time) continue;
        gyro_val = scale(add(head.gyro, tail.gyro), 0.5);
        acc_val = scale(add(head.acc, tail.acc), 0.5);
        acc_val = divs(scale(acc_val, 9.81), gravity_norm);
        if (head.timestamp < last_cloud_end_time) dt = tail.timestamp - last_cloud_end_time;
        else dt = tail.timestamp - head.timestamp;
        kf.predict(acc_val, gyro_val, dt, Q);
        last_gyro = sub(gyro_val, kf.x().bg);
        last_acc = add(mul(kf.x().rot, sub(acc_val, kf.x().ba)), kf.x().g);
        imu_poses_cache.push_back(Pose{tail.timestamp - cloud_time_begin, last_acc, last_gyro, kf.x().vel, kf.x().pos, kf.x().rot});
    }
    dt = cloud_time_end - imu_time_end;
    kf.predict(acc_val, gyro_val, dt, Q);
    last_imu = package.imus.back();
    last_cloud_end_time = cloud_time_end;

    const M3 cur_rot = kf.x().rot, cur_rot_ext = kf.x().rot_ext;
    const V3 cur_pos = kf.x().pos, cur_pos_ext = kf.x().pos_ext;
    if (package.size() == 0 || !compensate) return;
    CloudPoint* pts = package.pts();
    size_t ip = package.size() - 1;
    for (size_t kp = imu_poses_cache.size() - 1; kp != 0; kp--) {
        const Pose& head = imu_poses_cache[kp - 1];
        const Pose& tail = imu_poses_cache[kp];
        for (; pts[ip].curvature / double(1000) > head.offset; ip--) {
            dt = pts[ip].curvature / double(1000) - head.offset;
            const V3 point = v3(pts[ip].x, pts[ip].y, pts[ip].z);
            const M3 point_rot = mul(head.rot, so3_exp(scale(tail.gyro, dt)));
            const V3 point_pos = add(add(head.pos, scale(head.vel, dt)), scale(scale(scale(tail.acc, 0.5), dt), dt));
            const V3 inner = sub(add(mul(point_rot, add(mul(cur_rot_ext, point), cur_pos_ext)), point_pos), cur_pos);
            const V3 pc = mul(tr(cur_rot_ext), sub(mul(tr(cur_rot), inner), cur_pos_ext));
            pts[ip].x = (float)pc[0]; pts[ip].y = (float)pc[1]; pts[ip].z = (float)pc[2];
            if (ip == 0) break;
        }
    }
}

// lio_builder.cpp:65-116 without kf.predict: the averaged / rescaled IMU input, length and pose offset of every propagation step
// (the propagation itself and the pose list are then produced on the device, vmp_scan_raw_predict)
void LIOBuilder::collectImuSteps(SyncPackage& package, std::vector<vmp_imu_step>& steps) {
    imu_cache.clear();
    imu_cache.push_back(last_imu);
    imu_cache.insert(imu_cache.end(), package.imus.begin(), package.imus.end());
    const double imu_time_end = imu_cache.back().timestamp;
    const double cloud_time_begin = package.cloud_start_time, cloud_time_end = package.cloud_end_time;
    steps.clear();
    V3 acc_val = zeros<3, 1>(), gyro_val = zeros<3, 1>();
    double dt = 0.0;
    auto push = [&](double step_dt, double offset) {
        vmp_imu_step s;
        for (int c = 0; c < 3; c++) { s.acc[c] = acc_val[c]; s.gyro[c] = gyro_val[c]; }
        s.dt = step_dt; s.offset = offset;
        steps.push_back(s);
    };
    for (size_t i = 0; i + 1 < imu_cache.size(); i++) {
        const IMUData& head = imu_cache[i];
        const IMUData& tail = imu_cache[i + 1];
        if (tail.timestamp < last_cloud_end_time) continue;
        gyro_val = scale(add(head.gyro, tail.gyro), 0.5);
        acc_val = scale(add(head.acc, tail.acc), 0.5);
        acc_val = divs(scale(acc_val, 9.81), gravity_norm);
        if (head.timestamp < last_cloud_end_time) dt = tail.timestamp - last_cloud_end_time;
        else dt = tail.timestamp - head.timestamp;
        push(dt, tail.timestamp - cloud_time_begin);
    }
    push(cloud_time_end - imu_time_end, VMP_NO_POSE);
    last_imu = package.imus.back();
    last_cloud_end_time = cloud_time_end;
}

// the device path takes 2..64 poses (one per IMU sample of the scan)
bool LIOBuilder::imu_poses_fit(const SyncPackage& package) const { return package.imus.size() + 1 >= 2 && package.imus.size() + 1 <= 64; }

// lio_builder.cpp:175-248
int LIOBuilder::process(SyncPackage& package, vmp_scan_stats* stats) {
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (status == IMU_INIT) {
        if (initializeImu(package.imus)) { status = MAP_INIT; last_cloud_end_time = package.cloud_end_time; }
        return VMP_OK;
    }
    // every branch below stages the RAW cloud on the device (the first scan is never filtered, lio_builder.cpp:188-208): reject
    // a cloud that does not fit BEFORE the filter state, the IMU hand-over and the cloud itself are advanced
    if ((int)package.size() > config.max_points_per_scan) {
        set_error("LIOBuilder::process: %d points exceed max_points_per_scan=%d (it has to cover raw, unfiltered scans)", (int)package.size(), config.max_points_per_scan);
        return VMP_ERR_INVALID_ARG;
    }
    // MAP_INIT (once): everything on the host like the reference.  LIO_MAPPING: IMU propagation on the host, the point
    // loop of undistortCloud on the device in front of the update, one upload + one graph.
    const bool on_device = status == LIO_MAPPING && device_undistort && imu_poses_fit(package);
    if (on_device && device_predict) {
        // IMU propagation, motion compensation and update in one graph; the host only assembles the IMU input of the steps
        collectImuSteps(package, steps_);
        double last6[6];
        for (int c = 0; c < 3; c++) { last6[c] = last_acc[c]; last6[3 + c] = last_gyro[c]; }
        vmp_state xs;
        static_assert(sizeof(CloudPoint) == 16, "CloudPoint is x y z t");
        const int r = vmp_scan_raw_predict(map, &xs, kf.P(), reinterpret_cast<float*>(package.pts()), (int)package.size(), steps_.data(), (int)steps_.size(), Q,
                                           predict_started ? nullptr : last6, stats);
        if (r) return r;
        predict_started = true;
        kf.x() = st_load(reinterpret_cast<const double*>(&xs));
        return VMP_OK;
    }
    if (predict_started) { set_error("LIOBuilder::process: the IMU tail (last_acc / last_gyro) lives on the device once device_predict has run; it cannot be switched off mid-run"); return VMP_ERR_STATE; }
    undistortCloud(package, !on_device);
    const int n = (int)package.size();
    vmp_state xs;
    st_store(kf.x(), reinterpret_cast<double*>(&xs));
    prior_x = xs;
    std::memcpy(prior_P, kf.P(), sizeof(prior_P));
    if (on_device) {
        poses_.resize(imu_poses_cache.size());
        for (size_t k = 0; k < imu_poses_cache.size(); k++) {
            const Pose& p = imu_poses_cache[k];
            vmp_pose& q = poses_[k];
            q.offset = p.offset;
            for (int c = 0; c < 3; c++) { q.acc[c] = p.acc[c]; q.gyro[c] = p.gyro[c]; q.vel[c] = p.vel[c]; q.pos[c] = p.pos[c]; }
            for (int c = 0; c < 9; c++) q.rot[c] = p.rot.a[c];
        }
        static_assert(sizeof(CloudPoint) == 16, "CloudPoint is x y z t");
        const int r = vmp_scan_raw(map, &xs, kf.P(), reinterpret_cast<float*>(package.pts()), n, poses_.data(), (int)poses_.size(), stats);
        if (r) return r;
        kf.x() = st_load(reinterpret_cast<const double*>(&xs));
        return VMP_OK;
    }
    if (status == MAP_INIT) {
        xyz_.resize((size_t)n * 3);
        { const CloudPoint* c = package.pts(); for (int i = 0; i < n; i++) { xyz_[3 * i] = c[i].x; xyz_[3 * i + 1] = c[i].y; xyz_[3 * i + 2] = c[i].z; } }
        vmp_update_stats us;
        const int r = vmp_first_scan(map, &xs, kf.P(), xyz_.data(), n, &us);
        if (r) return r;
        if (stats) stats->map = us;
        status = LIO_MAPPING;
        return VMP_OK;
    }
    if (config.scan_resolution > 0.0) {     // lio_builder.cpp:215-219 after a host-side compensation: filter, then the update
        ds_.resize((size_t)n * 4);
        int m = 0;
        const int rd = vmp_downsample(map, reinterpret_cast<const float*>(package.pts()), n, config.scan_resolution, ds_.data(), n, &m);
        if (rd) return rd;
        xyz_.resize((size_t)m * 3);
        for (int i = 0; i < m; i++) { xyz_[3 * i] = ds_[4 * i]; xyz_[3 * i + 1] = ds_[4 * i + 1]; xyz_[3 * i + 2] = ds_[4 * i + 2]; }
        const int r = vmp_scan(map, &xs, kf.P(), xyz_.data(), m, stats);
        if (r) return r;
        kf.x() = st_load(reinterpret_cast<const double*>(&xs));
        return VMP_OK;
    }
    // lio_builder.cpp:224-229 reads x y z out of the PCL points; here they go straight into the pinned staging of the scan
    // host_ms of this branch covers the whole timed region lio_builder.cpp:224-246 as the caller sees it: reading the points out
    // of the caller's (pageable) cloud into the pinned staging, the upload, the graph, the results back in host memory
    const auto t_region = std::chrono::steady_clock::now();
    static_assert(sizeof(CloudPoint) == 16, "CloudPoint is x y z t");
    { const int rf = vmp_scan_buffer_fill(map, reinterpret_cast<const float*>(package.pts()), 4, n); if (rf) return rf; }      // (by the staging helpers: one core needs ~0.3 ms for 200 000 points)
    const int r = vmp_scan_staged(map, &xs, kf.P(), n, stats);           // posterior written back into xs / kf.P()
    if (stats) stats->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_region).count();
    if (r) return r;
    kf.x() = st_load(reinterpret_cast<const double*>(&xs));
    return VMP_OK;
}

}  // namespace vmp

struct vmp_lio_t { vmp::LIOBuilder b; };

extern "C" {

int vmp_lio_create(const vmp_config* cfg, vmp_lio* out) {
    if (!cfg || !out) { vmp::set_error("vmp_lio_create: null argument"); return VMP_ERR_INVALID_ARG; }
    vmp_lio_t* l = new vmp_lio_t();
    const int r = l->b.loadConfig(*cfg);
    if (r) { delete l; *out = nullptr; return r; }
    *out = l;
    return VMP_OK;
}
int vmp_lio_destroy(vmp_lio l) { delete l; return VMP_OK; }

int vmp_lio_process(vmp_lio l, const vmp_imu* imus, int n_imu, float* cloud_xyzc, int n, double t0, double t1, vmp_scan_stats* stats) {
    if (!l || (n_imu > 0 && !imus) || (n > 0 && !cloud_xyzc) || n_imu < 1) { vmp::set_error("vmp_lio_process: invalid argument"); return VMP_ERR_INVALID_ARG; }
    vmp::SyncPackage pk;
    pk.imus.resize((size_t)n_imu);
    for (int i = 0; i < n_imu; i++) {
        std::memcpy(pk.imus[i].acc.a, imus[i].acc, 24);
        std::memcpy(pk.imus[i].gyro.a, imus[i].gyro, 24);
        pk.imus[i].timestamp = imus[i].timestamp;
    }
    static_assert(sizeof(vmp::CloudPoint) == 4 * sizeof(float), "CloudPoint is x y z t");
    pk.ext_cloud = reinterpret_cast<vmp::CloudPoint*>(cloud_xyzc);        // process() edits the caller's cloud in place
    pk.ext_size = (size_t)n;
    pk.cloud_start_time = t0; pk.cloud_end_time = t1;
    return l->b.process(pk, stats);
}
int vmp_lio_state(vmp_lio l, vmp_state* x, double* P, int* status) {
    if (!l) return VMP_ERR_INVALID_ARG;
    if (x) vmp::st_store(l->b.kf.x(), reinterpret_cast<double*>(x));
    if (P) std::memcpy(P, l->b.kf.P(), sizeof(double) * 529);
    if (status) *status = (int)l->b.status;
    return VMP_OK;
}
vmp_handle vmp_lio_map(vmp_lio l) { return l ? l->b.map : nullptr; }
int vmp_lio_set_device_undistort(vmp_lio l, int on) {
    if (!l) return VMP_ERR_INVALID_ARG;
    l->b.device_undistort = on != 0;
    return VMP_OK;
}
int vmp_lio_set_cloud_writeback(vmp_lio l, int on) {
    if (!l) return VMP_ERR_INVALID_ARG;
    return vmp_set_raw_writeback(l->b.map, on);
}
int vmp_lio_set_device_predict(vmp_lio l, int on) {
    if (!l) return VMP_ERR_INVALID_ARG;
    l->b.device_predict = on != 0;
    return VMP_OK;
}
int vmp_lio_prior(vmp_lio l, vmp_state* x, double* P) {
    if (!l) return VMP_ERR_INVALID_ARG;
    if (x) *x = l->b.prior_x;
    if (P) std::memcpy(P, l->b.prior_P, sizeof(double) * 529);
    return VMP_OK;
}

}  // extern "C"
